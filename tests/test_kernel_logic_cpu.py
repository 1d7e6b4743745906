"""CPU: the SOURCE of csrc/rt_raycast.cu executed on host threads (tools/cuda_emu: one std::thread per CUDA thread, barriers
for __syncthreads and the warp collectives) against the oracle's brute-force definition, bit for bit.  Covers the measured
screen-space packet walk, the per-lane 3-D walk and the two experimental variants (refit passes, two-level region traversal
with and without overflow of its shared arrays).  Kernel logic only; the B200 parity suite is tests/*_gpu.py."""
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.timeout(600)
def test_raycast_kernels_on_the_cpu_emulator():
    sys.path.insert(0, os.path.join(ROOT, "tools", "cuda_emu"))
    import check_raycast
    assert check_raycast.main(600, 64, 32, quick=True) == 0


@pytest.mark.timeout(600)
def test_raycast_kernel_edge_cases_on_the_cpu_emulator():
    """camera inside the mesh (unbounded rectangles), single-triangle scene, sub-rectangle with a pitch, no cull rectangle"""
    sys.path.insert(0, os.path.join(ROOT, "tools", "cuda_emu"))
    import check_raycast
    assert check_raycast.edges(quick=True) == 0


@pytest.mark.timeout(600)
def test_raster_kernels_on_the_cpu_emulator():
    """raster_kernel / coverage_kernel / resolve_kernel / points, lesson08 + lesson09, clipping camera, composing draws"""
    sys.path.insert(0, os.path.join(ROOT, "tools", "cuda_emu"))
    import check_raster
    assert check_raster.main(900, 80, 48, quick=True) == 0


@pytest.mark.timeout(600)
def test_bvh_builders_on_the_cpu_emulator():
    """bounds / Morton / radix sort / Karras + refit / PLOC rounds: valid trees (soup, indexed, duplicate Morton codes, one
    triangle), and the emulated traversal over them reproduces the oracle's hits"""
    sys.path.insert(0, os.path.join(ROOT, "tools", "cuda_emu"))
    import check_bvh
    assert check_bvh.main(300, 48, 32, quick=True) == 0


@pytest.mark.timeout(600)
def test_gather_entry_points_on_the_cpu_emulator():
    """csrc/rt_peer.cu on the host: rt_copy_stripes (the tile partition's 3-D copy) against a row-by-row definition over random
    rects / stripes / rank counts, rt_push_tiles (tile-sparse push kernel) over cycles of frames, alloc / export / open / copy_rect"""
    sys.path.insert(0, os.path.join(ROOT, "tools", "cuda_emu"))
    import check_peer
    assert check_peer.main(quick=True) == 0
