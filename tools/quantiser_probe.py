"""Upsizing into uint16 / uint8 (expand kernel): how many values differ from the oracle, and by how much."""
import sys
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import numpy as np
import openimageio_b200 as oiio, oracle_py as oracle, synth
for st, dt in (("float", "uint16"), ("uint16", "uint16"), ("uint8", "uint8"), ("float", "uint8")):
    nd = {"float": np.float32, "uint16": np.uint16, "uint8": np.uint8}
    src = synth.image(300, 200, 3, nd[st], seed=77)
    ref = np.zeros((800, 1200, 3), nd[dt])
    oracle.resize(oracle.Img(ref), oracle.Img(src), "mitchell")
    got = np.zeros_like(ref)
    assert oiio.ImageBufAlgo.resize(oiio.ImageBuf(got), oiio.ImageBuf(src), filtername="mitchell")
    d = np.abs(got.astype(int) - ref.astype(int))
    print("%s -> %s: max diff %d LSB, differing values %.4f %%" % (st, dt, d.max(), 100.0 * (d > 0).mean()))
