"""tests/golden/make_host_math_golden.py -- TEST INFRASTRUCTURE.  Needs /root/reference (not available on the GPU box).

    python tests/golden/make_host_math_golden.py        ->  tests/golden/host_math_reference.npz

Runs the reference's OWN host transform helpers (rendering/_core.py:421-548: scale, translate, identity, rotate, perspective,
matmul, dot, to_array) on seeded inputs and stores inputs + outputs.  The reference imports pyopencl; oracle/clshim/fake_cl
stands in for it (no kernels run here).  `look_at` and `normalize` are NOT in the fixture: under NumPy 2 the reference's
make_float3(ndarray) path raises ("could not assign tuple of length 7 to structure with 4 fields"), so those two stay pinned
only by SURVEY.md appendix D's known answers (oracle/host_math.py).  NumPy version used is recorded in the file.
"""
import os
import sys

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
REFERENCE = os.environ.get("RENDERTOY_REFERENCE", "/root/reference")
sys.path.insert(0, REPO)
from oracle.clshim import fake_cl  # noqa: E402

fake_cl.install()
sys.path.insert(0, REFERENCE)
for k in [k for k in sys.modules if k == "rendering" or k.startswith("rendering.")]:
    del sys.modules[k]
import rendering as ref  # noqa: E402  -- the reference itself
assert ref.__file__.startswith(REFERENCE), ref.__file__


def flat(m):
    """float4x4 value / tuple -> 16 float32"""
    return np.asarray(m if not isinstance(m, tuple) else np.array(m, dtype=ref.float4x4)).reshape(-1).view(np.float32)[:16].copy()


def main():
    rng = np.random.default_rng(12)
    out = {"numpy_version": np.__version__}
    angles = np.concatenate([[0.0, 0.5, np.pi / 2, np.pi, -1.25, 6.0], rng.uniform(-10, 10, 26)])
    axes = np.concatenate([[[0, 1, 0], [1, 0, 0], [0, 0, 1]], rng.standard_normal((29, 3))]).astype(np.float32)
    axes[3:] /= np.linalg.norm(axes[3:], axis=1, keepdims=True)
    out["rotate_angle"], out["rotate_axis"] = angles, axes
    out["rotate"] = np.stack([flat(ref.rotate(float(a), ref.make_float3(*[float(x) for x in ax]))) for a, ax in zip(angles, axes)])
    out["rotate_f32_angle"] = np.stack([flat(ref.rotate(np.float32(a), ref.make_float3(*[float(x) for x in ax]))) for a, ax in zip(angles, axes)])
    s3 = rng.uniform(0.1, 4, (16, 3)).astype(np.float32)
    out["scale_in"], out["scale"] = s3, np.stack([flat(ref.scale(*[float(x) for x in s])) for s in s3])
    out["scale_uniform"] = np.stack([flat(ref.scale(float(s[0]))) for s in s3])
    t3 = rng.uniform(-3, 3, (16, 3)).astype(np.float32)
    out["translate_in"], out["translate"] = t3, np.stack([flat(ref.translate(*[float(x) for x in t])) for t in t3])
    out["translate_vec"] = np.stack([flat(ref.translate(ref.make_float3(*[float(x) for x in t]))) for t in t3])
    out["identity"] = flat(ref.identity())
    persp = [(np.pi / 4, 1.0, 0.01, 100.0), (np.pi / 4, 16 / 9, 0.01, 100.0), (1.0, 4 / 3, 0.1, 50.0), (0.3, 2.35, 1.0, 1000.0), (3.141593 / 4, 3840 / 2160, .01, 100.0)]
    out["perspective_in"] = np.array(persp)
    out["perspective"] = np.stack([flat(ref.perspective(f, a, n, fa)) for f, a, n, fa in persp])
    out["perspective_default_aspect"] = np.stack([flat(ref.perspective(aspect_ratio=w / h)) for w, h in ((512, 512), (1920, 1080), (3840, 2160), (333, 211))])
    ma, mb = rng.standard_normal((12, 16)).astype(np.float32), rng.standard_normal((12, 16)).astype(np.float32)
    out["matmul_a"], out["matmul_b"] = ma, mb
    out["matmul"] = np.stack([flat(ref.matmul(ref.make_float4x4(*[float(x) for x in a]), ref.make_float4x4(*[float(x) for x in b]))) for a, b in zip(ma, mb)])
    va = rng.standard_normal((12, 4)).astype(np.float32)
    out["matmul_vec_a"] = va
    out["matmul_vec"] = np.stack([np.asarray(np.array(ref.matmul(ref.make_float4(*[float(x) for x in v]), ref.make_float4x4(*[float(x) for x in b])), dtype=ref.float4)).reshape(-1).view(np.float32)[:4]
                                  for v, b in zip(va, mb)])
    d3a, d3b = rng.standard_normal((12, 3)).astype(np.float32), rng.standard_normal((12, 3)).astype(np.float32)
    out["dot_a"], out["dot_b"] = d3a, d3b
    out["dot"] = np.array([ref.dot(ref.make_float3(*[float(x) for x in a]), ref.make_float3(*[float(x) for x in b])) for a, b in zip(d3a, d3b)], np.float64)
    # the tutorial's World matrix: matmul(scale(1), rotate(t, y))
    ts = np.linspace(0, 6.2, 16)
    out["world_t"] = ts
    out["world"] = np.stack([flat(ref.matmul(ref.scale(1.0), ref.rotate(float(t), ref.make_float3(0, 1, 0)))) for t in ts])
    path = os.path.join(REPO, "tests", "golden", "host_math_reference.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, {k: np.asarray(v).shape for k, v in out.items() if k != "numpy_version"})


if __name__ == "__main__":
    main()
