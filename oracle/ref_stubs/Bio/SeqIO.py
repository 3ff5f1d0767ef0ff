"""import-time placeholder"""


def parse(*a, **k):
    raise NotImplementedError('Bio stub: SeqIO.parse')


def write(*a, **k):
    raise NotImplementedError('Bio stub: SeqIO.write')
