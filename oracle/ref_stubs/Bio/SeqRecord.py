"""import-time placeholder"""


class SeqRecord:
    def __init__(self, *a, **k):
        raise NotImplementedError('Bio stub: SeqRecord')
