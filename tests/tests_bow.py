from bow_tree import make_tree  # noqa: F401  (import shim: tests/ is on sys.path under pytest rootdir conftest)
