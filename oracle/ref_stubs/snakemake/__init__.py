"""import-time placeholder"""
from . import io
