"""nngp_b200 -- B200-native NNGP kernel construction + exact GP posterior inference.

Host-side mirror of the neural-tangents surface the reference (Kangfei/NNGP-src) calls on its hot
path: ``stax.serial/Dense/Relu``, ``batch``, ``predict.gradient_descent_mse_ensemble`` ->
``predict_fn``; all arithmetic runs in ``libnngp_b200.so`` (hand-written sm_100a CUDA behind the C ABI
of ``include/nngp_b200.h``).  No CPU fallback.
"""
__version__ = "0.1.0"
