"""``nt.batch`` mirror.  Reference: ``nt.batch(kernel_fn, device_count=0, batch_size=0)`` at train.py:166-168,
estimator.py:31-33, active/ActiveLearner.py:24-26 -- with those arguments neural-tangents only jit-wraps the
kernel function.  Here batching is internal to the CUDA path (row blocks sized to whole DMMA waves), so
this is the identity; the arguments are accepted and ignored."""


def batch(kernel_fn, batch_size: int = 0, device_count: int = -1, store_on_device: bool = True):
    return kernel_fn
