"""import-time placeholder"""


def use(*a, **k):
    pass


rcParams = {}
