"""Host-side logic of the multi-GPU paths, on CPU: band split, band + halo
source spans (checked by re-running the ORACLE on a source cropped to the
span), and a world_size-2 gloo run of the rank-side arithmetic."""
import os
import sys

import numpy as np
import pytest

import synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_band_split_covers_roi_exactly(oiio):
    for h in (1, 7, 2160, 4321):
        for n in (1, 2, 3, 8):
            roi = oiio.ROI(5, 105, 10, 10 + h, 0, 1, 0, 4)
            rows = []
            for i in range(n):
                b = oiio.band_split(roi, n, i)
                assert (b.xbegin, b.xend) == (5, 105)
                rows += list(range(b.ybegin, b.yend))
            assert rows == list(range(10, 10 + h))


@pytest.mark.parametrize("filt,shape", [("lanczos3", (120, 160)),
                                        ("blackman-harris", (400, 300)),
                                        ("box", (60, 80)),
                                        ("radial-lanczos3", (100, 90))])
def test_source_rows_are_sufficient_and_tight(oiio, oracle, filt, shape):
    """Resizing a band from ONLY the rows b2r_source_rows names (source data
    window cropped to them) must equal resizing it from the whole source."""
    src = synth.image(200, 240, 3, np.float32, seed=77)
    dh, dw = shape
    full = np.zeros((dh, dw, 3), np.float32)
    oracle.resize(oracle.Img(full), oracle.Img(src), filt)
    S = oiio.ImageBuf(src)
    D = oiio.ImageBuf(np.zeros((dh, dw, 3), np.float32))
    n = 5
    for i in range(n):
        band = oiio.band_split(D.roi, n, i)
        r0, r1 = oiio.source_rows(D, S, filt, roi=band)
        assert 0 <= r0 < r1 <= 240
        crop = np.ascontiguousarray(src[r0:r1])
        out = np.zeros_like(full)
        oracle.resize(oracle.Img(out), oracle.Img(crop, 0, r0, (0, 0, 200, 240)),
                      filt, roi=(0, dw, band.ybegin, band.yend))
        assert np.array_equal(out[band.ybegin:band.yend],
                              full[band.ybegin:band.yend]), (i, r0, r1)
        # tight: the halo is the filter radius (+ the 8-row margin either side
        # that covers the kernels' rounded-up row windows), not the whole image
        if filt == "lanczos3" and 0 < i < n - 1:
            ratio = 240 / dh
            assert (r1 - r0) <= (band.yend - band.ybegin) * ratio \
                + 2 * 3 * max(1, ratio) + 4 + 16


def test_shard_frames(oiio):
    for n in (1, 7, 256):
        for w in (1, 2, 4, 8):
            got = sorted(sum((oiio.shard_frames(n, w, r) for r in range(w)), []))
            assert got == list(range(n))


_WORKER = r'''
import os, sys
sys.path.insert(0, sys.argv[1]); sys.path.insert(0, os.path.join(sys.argv[1], "tests"))
import numpy as np, torch, torch.distributed as dist
import openimageio_b200 as oiio, oracle_py, synth
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
src = synth.image(96, 128, 4, np.float32, seed=5)      # same on every rank
dh, dw = 64, 48
S = oiio.ImageBuf(src); D = oiio.ImageBuf(np.zeros((dh, dw, 4), np.float32))
band = oiio.band_split(D.roi, world, rank)
r0, r1 = oiio.source_rows(D, S, "lanczos3", roi=band)
# each rank resizes its band from its own band+halo rows only (oracle = CPU)
crop = np.ascontiguousarray(src[r0:r1])
out = np.zeros((dh, dw, 4), np.float32)
oracle_py.resize(oracle_py.Img(out), oracle_py.Img(crop, 0, r0, (0, 0, 96, 128)),
                 "lanczos3", roi=(0, dw, band.ybegin, band.yend))
# bands are disjoint: a sum-reduce assembles the image (test plumbing only)
t = torch.from_numpy(out); dist.all_reduce(t)
spans = [None] * world
dist.all_gather_object(spans, (band.ybegin, band.yend, r0, r1))
frames = [None] * world
dist.all_gather_object(frames, oiio.shard_frames(10, world, rank))
if rank == 0:
    ref = np.zeros((dh, dw, 4), np.float32)
    oracle_py.resize(oracle_py.Img(ref), oracle_py.Img(src), "lanczos3")
    assert np.array_equal(t.numpy(), ref)
    assert [s[0] for s in spans] == [0, 32] and [s[1] for s in spans] == [32, 64]
    assert spans[0][2] == 0 and spans[1][3] == 128 and spans[0][3] > spans[1][2]
    assert sorted(sum(frames, [])) == list(range(10))
    print("OK")
dist.destroy_process_group()
'''


def test_world_size_2_gloo(tmp_path):
    import subprocess
    w = tmp_path / "worker.py"
    w.write_text(_WORKER)
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    out = subprocess.run(
        [sys.executable, "-m", "torch.distributed.run", "--nnodes=1",
         "--nproc-per-node=2", "--master-addr", "127.0.0.1", "--master-port",
         "29611", str(w), ROOT],
        capture_output=True, text=True, timeout=600, env=env)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    assert "OK" in out.stdout


def test_source_rows_with_overscan_and_crop(oiio, oracle):
    """Data window larger than (overscan) or smaller than (crop) the display
    window: the staged span must still cover every tap after WrapClamp."""
    for (sx, sy, full) in [(-8, -6, (0, 0, 64, 64)), (20, 30, (0, 0, 128, 128))]:
        src = synth.image(80, 76, 3, np.float32, seed=78)
        dh, dw = 40, 36
        dfull = (0, 0, dw, dh)
        ref = np.zeros((dh, dw, 3), np.float32)
        oracle.resize(oracle.Img(ref, 0, 0, dfull), oracle.Img(src, sx, sy, full))
        S = oiio.ImageBuf(src)
        sp = S.spec()
        sp.x, sp.y = sx, sy
        sp.full_x, sp.full_y, sp.full_width, sp.full_height = full
        D = oiio.ImageBuf(np.zeros((dh, dw, 3), np.float32))
        for i in range(4):
            band = oiio.band_split(D.roi, 4, i)
            r0, r1 = oiio.source_rows(D, S, "", roi=band)
            crop = np.ascontiguousarray(src[r0:r1])
            out = np.zeros_like(ref)
            oracle.resize(oracle.Img(out, 0, 0, dfull),
                          oracle.Img(crop, sx, sy + r0, full), "",
                          roi=(0, dw, band.ybegin, band.yend))
            assert np.array_equal(out[band.ybegin:band.yend],
                                  ref[band.ybegin:band.yend]), (sx, i)


_MIP_WORKER = r'''
import os, sys
sys.path.insert(0, sys.argv[1]); sys.path.insert(0, os.path.join(sys.argv[1], "tests"))
import numpy as np, torch, torch.distributed as dist
import oracle_py, synth
from openimageio_b200 import mipdist
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()

def level_fn(dst_t, b0, sw, sh, src_t, r0, w, h, filt):
    d, s = dst_t.numpy(), src_t.numpy()
    if filt == "box":
        assert 2 * sw == w and r0 == 2 * b0          # aligned 2x2 case only here
        oracle_py.resize_block(oracle_py.Img(d), oracle_py.Img(s))
    else:
        oracle_py.resize(oracle_py.Img(d, 0, b0, (0, 0, sw, sh)),
                         oracle_py.Img(s, 0, r0, (0, 0, w, h)), filt,
                         roi=(0, sw, b0, b0 + d.shape[0]))

def tail_fn(full_t, filt):
    return [torch.from_numpy(a) for a in oracle_py.mip_chain(full_t.numpy(), filt)]

ok = True
for (W, H, filt) in [(96, 160, "lanczos3"), (64, 128, "box"), (80, 144, "triangle")]:
    lvl0 = synth.image(W, H, 3, np.float32, seed=9)            # same on every rank
    y0, y1 = mipdist.split_rows(H, world, rank)
    chain = mipdist.DistributedMipChain(torch.from_numpy(lvl0[y0:y1].copy()), W, H, filt,
                                        level_fn=level_fn, tail_fn=tail_fn,
                                        min_band_rows=4)
    ref = [lvl0] + oracle_py.mip_chain(lvl0, filt)
    assert chain.nlevels == len(ref), (chain.nlevels, len(ref))
    nb = len(chain.bands)
    assert nb >= 3, nb                                        # really ran banded levels
    for k in range(1, nb):
        full = chain.gather_level(k)
        if rank == 0:
            # a banded level is computed from the SAME previous level as the
            # reference chain, band by band: identical arithmetic
            assert np.array_equal(full.numpy(), ref[k]), (filt, k)
    if rank == 0:
        for i, t in enumerate(chain.tail):
            assert np.array_equal(t.numpy(), ref[nb + i]), (filt, nb + i)
    if filt != "box":
        assert chain.halo_bytes > 0                           # halos did travel
if rank == 0:
    print("OK")
dist.destroy_process_group()
'''


def test_distributed_mip_chain_gloo(tmp_path):
    """The banded MIP chain's halo exchange and gather-to-rank-0 logic with
    the oracle as the compute step, world_size 2 over gloo."""
    import subprocess
    w = tmp_path / "mipworker.py"
    w.write_text(_MIP_WORKER)
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    out = subprocess.run(
        [sys.executable, "-m", "torch.distributed.run", "--nnodes=1",
         "--nproc-per-node=2", "--master-addr", "127.0.0.1", "--master-port",
         "29612", str(w), ROOT],
        capture_output=True, text=True, timeout=900, env=env)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]
    assert "OK" in out.stdout
