class _Backend:
    platform = "cpu"   # train.py:288 only uses this to set args.cuda; the JAX side stays "cpu" by construction


def get_backend(*_a, **_k):
    return _Backend()
