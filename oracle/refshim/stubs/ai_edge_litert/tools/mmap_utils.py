"""No-op stand-in for ai_edge_litert.tools.mmap_utils (test infrastructure)."""


def advise_dont_need(*args, **kwargs):
  del args, kwargs


def get_file_contents(path):
  with open(path, "rb") as f:
    return f.read()


def set_file_contents(path, data):
  with open(path, "wb") as f:
    f.write(data)
