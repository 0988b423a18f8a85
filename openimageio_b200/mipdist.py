"""Row-band MIP chain across GPUs: one process per GPU (`torch.distributed`).

The level loop of `write_mipmap` (src/libOpenImageIO/maketexture.cpp:823-986)
with every level split into full-width row bands, one band per rank (the
reference's own `parallel_image` split, imagebufalgo_util.h:37-80).  Level k+1's
band on rank g needs level k's rows of that band plus the filter-radius halo
(xform.cpp:626-633: <= 6 rows for lanczos3, none for the 2x2 box): the halo
rows that live on neighbouring ranks are fetched with point-to-point
send/recv (NCCL over NVLink P2P on a B200 box) -- a neighbour exchange, not a
collective.  Once bands get shorter than the halo the remaining levels are
gathered onto rank 0 and finished there with the single-GPU chain.

The compute step and the row-span arithmetic are injectable so the exchange
logic is testable on CPU (gloo) with a CPU stand-in for the kernels supplied
by the test (tests/test_sharding_cpu.py); the defaults call the C ABI (`b2r_mip_level`,
`b2r_source_rows`).
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import (ImageBuf, ROI, B2RError, _Roi, _options, _current_stream, lib,
               source_rows)


def split_rows(h, world, rank):
    """Band `rank` of `world` over h rows (b2r_band_split's rule)."""
    return h * rank // world, h * (rank + 1) // world


def default_rows_fn(filtername, w, h, sw, sh, b0, b1):
    """Rows [r0, r1) of the w x h level that rows [b0, b1) of the sw x sh level
    read."""
    if b1 <= b0:
        return 0, 0
    if filtername == "box":
        if 2 * sw == w and w % 2 == 0 and h % 2 == 0:
            return 2 * b0, 2 * b1          # resize_block_2pass: rows 2y, 2y+1
        # bilinear at NDC centres: floor(y') and floor(y')+1, plus slack
        r0 = max(0, (b0 * h) // sh - 2)
        r1 = min(h, -((-b1 * h) // sh) + 2)
        return r0, r1
    dst = ImageBuf(np.empty((1, 1, 1), np.float32))
    src = ImageBuf(np.empty((1, 1, 1), np.float32))
    d, s = dst.spec(), src.spec()
    d.width = d.full_width = sw
    d.height = d.full_height = sh
    s.width = s.full_width = w
    s.height = s.full_height = h
    return source_rows(dst, src, filtername, roi=ROI(0, sw, b0, b1, 0, 1, 0, 1))


def default_level_fn(dst_t, b0, sw, sh, src_t, r0, w, h, filtername):
    """One band of one level on this rank's GPU through the C ABI."""
    c = dst_t.shape[2]
    D, S = ImageBuf(dst_t), ImageBuf(src_t)
    ds, ss = D.spec(), S.spec()
    ds.y, ds.full_x, ds.full_y, ds.full_width, ds.full_height = b0, 0, 0, sw, sh
    ss.y, ss.full_x, ss.full_y, ss.full_width, ss.full_height = r0, 0, 0, w, h
    d, s = D._c(), S._c()
    r = _Roi(0, sw, b0, b0 + dst_t.shape[0], 0, c)
    o = _options(dst_t.device.index or 0, 0, _current_stream(D))
    err = C.create_string_buffer(512)
    rc = lib().b2r_mip_level(C.byref(d), C.byref(s), filtername.encode(),
                             C.byref(r), C.byref(o), err, 512)
    if rc:
        raise B2RError(err.value.decode())


class DistributedMipChain:
    """bands[k] = (y0, y1, tensor) : this rank's rows of level k (None once
    the chain has been gathered onto rank 0; rank 0 then holds `tail`, the
    remaining full levels)."""

    def __init__(self, band0, full_w, full_h, filtername="box", group=None,
                 level_fn=default_level_fn, rows_fn=default_rows_fn,
                 min_band_rows=8, tail_fn=None):
        import torch
        import torch.distributed as dist
        self.dist, self.group = dist, group
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)
        self.filtername = filtername
        self._level_fn = level_fn
        y0, y1 = split_rows(full_h, self.world, self.rank)
        assert band0.shape[0] == y1 - y0 and band0.shape[1] == full_w
        self.sizes = [(full_w, full_h)]
        self.bands = [(y0, y1, band0)]
        self.tail = None
        self.halo_bytes = 0
        w, h = full_w, full_h
        while w > 1 or h > 1:
            sw, sh = (w // 2 if w > 1 else w), (h // 2 if h > 1 else h)
            spans = [rows_fn(filtername, w, h, sw, sh, *split_rows(sh, self.world, q))
                     for q in range(self.world)]
            owners = [split_rows(h, self.world, q) for q in range(self.world)]
            # stay banded while every rank has a real band and each halo lives
            # on the rank itself or its direct neighbours
            ok = sh >= self.world * min_band_rows
            for q in range(self.world):
                r0, r1 = spans[q]
                lo = owners[max(q - 1, 0)][0]
                hi = owners[min(q + 1, self.world - 1)][1]
                ok = ok and r0 >= lo and r1 <= hi
            if not ok:
                self._gather_and_finish(w, h, tail_fn)
                return
            b0, b1 = split_rows(sh, self.world, self.rank)
            r0, r1 = spans[self.rank]
            s0, s1, own = self.bands[-1]
            src = torch.empty((r1 - r0, w, own.shape[2]), dtype=own.dtype,
                              device=own.device)
            lo, hi = max(r0, s0), min(r1, s1)
            if hi > lo:
                src[lo - r0:hi - r0].copy_(own[lo - s0:hi - s0])
            ops, keep = [], []
            for q in (self.rank - 1, self.rank + 1):
                if q < 0 or q >= self.world:
                    continue
                q0, q1 = owners[q]
                # rows of q's band that I need
                a, b = max(r0, q0), min(r1, q1)
                if b > a:
                    ops.append(dist.P2POp(dist.irecv, src[a - r0:b - r0], q,
                                          group=group))
                    self.halo_bytes += (b - a) * w * own.shape[2] * own.element_size()
                # rows of my band that q needs
                a, b = max(spans[q][0], s0), min(spans[q][1], s1)
                if b > a:
                    t = own[a - s0:b - s0].contiguous()
                    keep.append(t)
                    ops.append(dist.P2POp(dist.isend, t, q, group=group))
            if ops:
                for req in dist.batch_isend_irecv(ops):
                    req.wait()
            dst = torch.empty((b1 - b0, sw, own.shape[2]), dtype=own.dtype,
                              device=own.device)
            level_fn(dst, b0, sw, sh, src, r0, w, h, filtername)
            self.sizes.append((sw, sh))
            self.bands.append((b0, b1, dst))
            w, h = sw, sh

    # ------------------------------------------------------------------
    def _gather_and_finish(self, w, h, tail_fn):
        import torch
        dist = self.dist
        s0, s1, own = self.bands[-1]
        full = self.gather_level(len(self.bands) - 1)
        if self.rank == 0:
            if tail_fn is not None:
                self.tail = tail_fn(full, self.filtername)
            else:
                # remaining levels stay on rank 0's GPU: the same per-level
                # ABI call on whole images, no host round trip
                self.tail = []
                cur, cw, ch = full, w, h
                while cw > 1 or ch > 1:
                    nw, nh = (cw // 2 if cw > 1 else cw), (ch // 2 if ch > 1 else ch)
                    nxt = torch.empty((nh, nw, cur.shape[2]), dtype=cur.dtype,
                                      device=cur.device)
                    self._level_fn(nxt, 0, nw, nh, cur, 0, cw, ch, self.filtername)
                    self.tail.append(nxt)
                    cur, cw, ch = nxt, nw, nh
            ww, hh = w, h
            for t in self.tail:
                self.sizes.append((t.shape[1], t.shape[0]))
        else:
            ww, hh = w, h
            while ww > 1 or hh > 1:
                ww, hh = (ww // 2 if ww > 1 else ww), (hh // 2 if hh > 1 else hh)
                self.sizes.append((ww, hh))

    def gather_level(self, k):
        """Assemble banded level k on rank 0 (None elsewhere)."""
        import torch
        dist = self.dist
        y0, y1, own = self.bands[k]
        w, h = self.sizes[k]
        if self.rank == 0:
            full = torch.empty((h, w, own.shape[2]), dtype=own.dtype,
                               device=own.device)
            full[y0:y1].copy_(own)
            reqs = []
            for q in range(1, self.world):
                q0, q1 = split_rows(h, self.world, q)
                if q1 > q0:
                    reqs.append(dist.irecv(full[q0:q1], q, group=self.group))
            for r in reqs:
                r.wait()
            return full
        if y1 > y0:
            dist.send(own.contiguous(), 0, group=self.group)
        return None

    @property
    def nlevels(self):
        return len(self.sizes)
