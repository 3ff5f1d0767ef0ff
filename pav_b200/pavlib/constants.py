"""Constants shared with the reference (pavlib/constants.py:55)."""
ERR_INV_FAIL = 125
