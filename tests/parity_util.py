"""Shared helpers for parity tests: run two worlds side by side and compare state + contact tables."""
import numpy as np

TABLES = [("pair_keys", np.uint64), ("pair_nsc", np.int32), ("pair_npts", np.int32), ("pair_color", np.int32),
          ("pair_normal", np.float32), ("pair_points", np.uint32), ("pair_data", np.float32), ("pair_sc", np.uint32)]


def compare_worlds(a, b, tables=True):
    """Returns a dict of max abs differences (0.0 everywhere = bit-exact) between two worlds exposing
    body_states() and debug_read()."""
    pa, va = a.body_states()
    pb, vb = b.body_states()
    out = {"pose": float(np.abs(pa - pb).max()) if pa.size else 0.0, "vel": float(np.abs(va - vb).max()) if va.size else 0.0,
           "pose_bits": int((pa.view(np.uint32) != pb.view(np.uint32)).sum()),
           "vel_bits": int((va.view(np.uint32) != vb.view(np.uint32)).sum())}
    if tables:
        for name, dt in TABLES:
            ta, tb = a.debug_read(name, dt), b.debug_read(name, dt)
            if ta.shape != tb.shape:
                out[name] = f"shape {ta.shape} vs {tb.shape}"
            else:
                out[name] = int((ta != tb).sum())
    return out


def is_exact(d):
    return all(v == 0 for v in d.values())
