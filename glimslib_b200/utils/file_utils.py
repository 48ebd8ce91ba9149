"""``ensure_dir_exists`` (counterpart of ``glimslib/utils/file_utils.py``)."""
import os


def ensure_dir_exists(path):
    d = path if not os.path.splitext(path)[1] else os.path.dirname(path)
    if d:
        os.makedirs(d, exist_ok=True)
    return d
