"""import-time placeholder"""


def venn2(*a, **k):
    raise NotImplementedError


def venn3(*a, **k):
    raise NotImplementedError
