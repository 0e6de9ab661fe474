"""The N > 1 path with the CUDA backend, launched under torch.distributed.run as the product is: every rank runs
slabs.segment_tile / raster_tile through the C ABI on its chunk, and the stitched result must be the UNDIVIDED tile's --
labels, planeIdx and plane count of seg_plane::get_planes over the whole cloud (my_function.cpp:180-258), kNN rows and
normals of every owned point, and the raster's doubles and bytes (TMC3.cpp:81-198) -- bit for bit against the oracle.

  * gloo, two and three ranks sharing GPU 0: runs on a one-GPU box (collectives staged through the host);
  * nccl, one GPU per rank: needs two GPUs, skipped otherwise (the driver's scaling run covers it at 2/4/8)."""
import os
import socket
import subprocess
import sys

import numpy as np
import pytest

import cases
import oracle_lib as O

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _launch(world, backend, case, n, halo, out):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", str(_free_port()), os.path.join(HERE, "gpu_tile_worker.py"), "--backend", backend, "--case", case,
           "--points", str(n), "--halo", str(halo), "--out", str(out)]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    return [np.load(os.path.join(out, f"r{k}.npz")) for k in range(world)]


def _check(res, xyz):
    P = O.pipeline(xyz)
    g = P["grow"]
    seen = np.zeros(len(xyz), bool)
    for r in res:
        c0, c1 = int(r["c0"]), int(r["c1"])
        assert int(r["n_planes"]) == g.n_planes
        assert np.array_equal(r["labels"], g.label[c0:c1])
        assert np.array_equal(r["plane_idx"], g.plane_idx[c0:c1])
        og = r["owned_gid"]
        seen[og] = True
        assert np.array_equal(r["owned_rows"], P["neigh"][og])
        assert np.array_equal(r["owned_normals"].view(np.int64), P["normals"][og].view(np.int64))
        assert int(r["launches"]) > 0  # the CUDA kernels of this rank's context ran
    assert seen.all()
    W, H = int(P["wh"][0]), int(P["wh"][1])
    ref = O.raster(P["xyz"], P["mx"][2] - P["mn"][2], W, H)
    oa, ob, _, _ = O.save_image(ref)
    img = np.concatenate([r["image"] for r in res], axis=1)
    assert (int(res[0]["W"]), int(res[0]["H"])) == (W, H)
    assert np.array_equal(img.view(np.int64), ref.view(np.int64))
    assert np.array_equal(np.concatenate([r["a"] for r in res], axis=1), oa)
    assert np.array_equal(np.concatenate([r["b"] for r in res], axis=1), ob)
    return g


@pytest.mark.parametrize("world,case,n,halo", [(2, "block", 200000, 400), (3, "block", 150000, 100), (2, "quantised", 60000, 400)])
def test_tile_on_one_gpu_is_the_undivided_tile(world, case, n, halo, tmp_path):
    res = _launch(world, "gloo", case, n, halo, tmp_path)
    g = _check(res, getattr(cases, case)(n=n))
    assert g.n_planes > 0 and sum(int(r["n_halo"]) for r in res) > 0
    if halo == 100:
        assert int(res[0]["halo"]) > 100  # the sufficiency loop ran


def test_tile_over_nccl_is_the_undivided_tile(tmp_path):
    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs (the one-GPU variant above runs the same host logic over gloo)")
    world = min(4, torch.cuda.device_count())
    res = _launch(world, "nccl", "block", 400000, 400, tmp_path)
    _check(res, cases.block(n=400000))
