"""Variant-ID versioning, host-side string work (reference: dep/svpop/svpoplib/variant.py:664-752)."""
import collections
import re

import pandas as pd

_VERSIONED = re.compile(r'.*\.\d+$')


def version_id_name(name, id_set):
    """Next free ``name.N`` given the IDs already taken."""
    if name not in id_set:
        return name
    if _VERSIONED.match(name):
        stem, ver = name.rsplit('.', 1)
        ver = int(ver) + 1
    else:
        stem, ver = name, 1
    cand = f'{stem}.{ver}'
    while cand in id_set:
        ver += 1
        cand = f'{stem}.{ver}'
    return cand


def version_id(id_col, existing_id_set=None):
    """De-duplicate a Series of IDs by appending ``.N``; the first occurrence keeps the bare ID."""
    counts = collections.Counter()
    if existing_id_set is not None:
        counts.update(existing_id_set)
    values = id_col.tolist()
    counts.update(values)
    dup = {k for k, c in counts.items() if c > 1}
    if not dup:
        return id_col
    taken = set(values) - dup
    if existing_id_set is not None:
        taken |= set(existing_id_set)
    for i, name in enumerate(values):
        if name in dup:
            new = version_id_name(name, taken)
            values[i] = new
            taken.add(new)
    return pd.Series(values, index=id_col.index, dtype=object, name=id_col.name)
