"""Bio.Seq.Seq(s).reverse_complement() with Biopython's IUPAC ambiguous-DNA complement table,
case preserved (used at pavlib/cigarcall.py:70 and pavlib/seq.py:355-358)."""

_COMP = str.maketrans(
    'ACGTMRWSYKVHDBNacgtmrwsykvhdbn',
    'TGCAKYWSRMBDHVNtgcakywsrmbdhvn',
)


class Seq:
    def __init__(self, data):
        self._data = str(data)

    def reverse_complement(self):
        return Seq(self._data.translate(_COMP)[::-1])

    def complement(self):
        return Seq(self._data.translate(_COMP))

    def upper(self):
        return Seq(self._data.upper())

    def __str__(self):
        return self._data

    def __len__(self):
        return len(self._data)

    def __getitem__(self, k):
        r = self._data[k]
        return Seq(r) if isinstance(k, slice) else r
