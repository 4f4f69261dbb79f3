"""Deterministic synthetic tensors shared by the golden generator, the oracle tests and the
GPU parity tests.  Pure integer hashing (no RNG state, no torch/numpy version dependence), and
every value is exactly representable in bf16, so the bf16 tensor-core path sees the same
weights / inputs as the fp32 oracle."""
import numpy as np

_M1 = np.uint64(0x9E3779B97F4A7C15)
_M2 = np.uint64(0xBF58476D1CE4E5B9)
_M3 = np.uint64(0x94D049BB133111EB)


def _mix(idx, seed):
    with np.errstate(over="ignore"):
        z = idx.astype(np.uint64) * _M1 + np.uint64(seed) * _M2 + _M3
        z ^= z >> np.uint64(30)
        z *= _M2
        z ^= z >> np.uint64(27)
        z *= _M3
        z ^= z >> np.uint64(31)
    return z


def bf16_round(a):
    """Round-to-nearest-even to bf16, returned as float32."""
    a = np.ascontiguousarray(a, dtype=np.float32)
    u = a.view(np.uint32).astype(np.uint64)
    u = (u + np.uint64(0x7FFF) + ((u >> np.uint64(16)) & np.uint64(1))) & np.uint64(0xFFFF0000)
    return u.astype(np.uint32).view(np.float32).reshape(a.shape)


def det_uniform(shape, seed, lo=-1.0, hi=1.0, bf16=True):
    """Uniform in [lo, hi), deterministic in (shape, seed)."""
    n = int(np.prod(shape))
    z = _mix(np.arange(n, dtype=np.uint64), seed)
    u = (z >> np.uint64(40)).astype(np.float64) / float(1 << 24)
    a = (lo + (hi - lo) * u).astype(np.float32).reshape(shape)
    return bf16_round(a) if bf16 else a


def name_seed(name, base):
    h = 1469598103934665603
    for ch in name.encode():
        h = ((h ^ ch) * 1099511628211) & 0xFFFFFFFFFFFFFFFF
    return (h ^ (base * 0x9E3779B97F4A7C15)) & 0x7FFFFFFFFFFFFFFF


def det_state_dict(module, base_seed):
    """Deterministic parameters for every entry of ``module.state_dict()``:
    conv / linear weights ~ U(-a, a) with a = sqrt(3 / fan_in) (unit-gain variance),
    norm weight in [0.6, 1.4], norm bias and running_mean in [-0.2, 0.2], running_var in
    [0.6, 1.4]; ``num_batches_tracked`` left untouched."""
    import torch
    out = {}
    for k, v in module.state_dict().items():
        if k.endswith("num_batches_tracked"):
            out[k] = v.clone()
            continue
        s = name_seed(k, base_seed)
        shp = tuple(v.shape)
        if v.ndim >= 2:
            # Conv: [Cout, Cin, k..]; ConvTranspose: [Cin, Cout, k..].  fan_in ~ numel / shape[0]
            # (for the k3/s2 transposed convs only ~1/8 of the taps hit an output, so this is
            # conservative either way).
            fan_in = max(1, int(np.prod(shp[1:])))
            a = float(np.sqrt(3.0 / fan_in))
            arr = det_uniform(shp, s, -a, a)
        elif k.endswith("running_var") or k.endswith("weight"):
            arr = det_uniform(shp, s, 0.6, 1.4)
        else:
            arr = det_uniform(shp, s, -0.2, 0.2)
        out[k] = torch.from_numpy(arr.copy())
    return out
