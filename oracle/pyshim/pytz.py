"""Minimal stub of `pytz` (absent here; reference config.py:5 imports it only for a log-file name)."""
import datetime


def timezone(name):
    return datetime.timezone(datetime.timedelta(hours=8), name)
