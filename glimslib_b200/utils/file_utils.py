"""``ensure_dir_exists`` / ``get_file_extension`` (counterpart of ``glimslib/utils/file_utils.py:5-40``)."""
import os


def get_file_extension(path_to_file):
    """Text after the last '.' of the file name, or None when the name has no '.'."""
    name = os.path.split(path_to_file)[-1]
    return name.rsplit(".", 1)[1] if "." in name else None


def ensure_dir_exists(path):
    """Creates the directory ``path`` names: ``path`` itself when its last component has no extension, else its parent."""
    d = path if get_file_extension(path) is None else os.path.dirname(path)
    if d:
        os.makedirs(d, exist_ok=True)
    return d
