"""The ctypes boundary without a GPU: every tensor-level wrapper of `fasterseg_b200.functional` is driven with CPU tensors
against a FAKE library whose entry points are `ctypes.CFUNCTYPE` callbacks built from the same signature table the real
binding uses (`_lib._SIGS`).  ctypes therefore converts / type-checks every argument exactly as it would for libfsb200.so
(wrong arity, a float where the header has an int, a tensor instead of a pointer -> the call raises), and the callbacks
record what arrived, so the tests can assert that descriptors, pointers, channel strides, counts and flags are the ones the
tensors imply.  No arithmetic happens; outputs are whatever `torch.empty` returned."""
import ctypes as C

import pytest
import torch

from fasterseg_b200 import _lib
from fasterseg_b200 import functional as F_
from tests import cpu_backend


class FakeLib:
    def __init__(self):
        self.calls = []
        self._keep = []
        for name, (res, args) in _lib._SIGS.items():
            proto = C.CFUNCTYPE(res, *args)

            def recorder(*a, _name=name, _res=res):
                vals = []
                for v in a:
                    if isinstance(v, C.POINTER(_lib.ConvDesc)):
                        d = v.contents
                        vals.append({f: getattr(d, f) for f, _ in _lib.ConvDesc._fields_})
                    else:
                        vals.append(v)
                self.calls.append((_name, vals))
                if _name in ("fsb_conv_packed_bytes", "fsb_conv_packed_dgrad_bytes"):
                    return 4096
                if _name == "fsb_abi_version":
                    return _lib.ABI_VERSION
                if _name in ("fsb_conv_stats_rows", "fsb_stat_rows", "fsb_wsum_rows"):
                    return 3       # partial statistic rows the fake "kernels" write
                return None if _res is C.c_char_p else 0

            cb = proto(recorder)
            self._keep.append(cb)
            setattr(self, name, cb)

    def last(self, name):
        for n, vals in reversed(self.calls):
            if n == name:
                return vals
        raise AssertionError("%s was not called; calls: %s" % (name, [n for n, _ in self.calls]))


@pytest.fixture
def fake(monkeypatch):
    lib = FakeLib()
    monkeypatch.setattr(_lib, "lib", lambda: lib)
    monkeypatch.setattr(F_, "_stream", lambda: 0)
    monkeypatch.setattr(F_, "_on_device", lambda t: True)         # the product insists on CUDA tensors; everything else is real
    return lib


def act(N, C_, H, W, dtype=torch.float16, wide=0):
    """NHWC activation view; `wide` extra channels make it a slice of a wider (concat) buffer"""
    t = F_.empty_nhwc(N, C_ + wide, H, W, "cpu", dtype=dtype)
    return t[:, :C_] if wide else t


def test_empty_nhwc_layout():
    for C_ in (19, 32, 1, 8):
        t = F_.empty_nhwc(2, C_, 5, 7, "cpu")
        N, Cc, H, W, cs = cpu_backend.nhwc_info(t)
        assert (N, Cc, H, W) == (2, C_, 5, 7) and cs == (C_ + 7) // 8 * 8 and t.dtype == torch.float16
        assert t.stride() == (5 * 7 * cs, 1, 7 * cs, cs)
        # the padded lanes of the LAST pixel belong to the allocation too (vector stores touch them)
        assert t.untyped_storage().nbytes() >= 2 * 5 * 7 * cs * 2


def test_conv_fwd_marshalling(fake):
    x = act(2, 32, 12, 20, wide=16)
    wp = torch.empty(2048, dtype=torch.float16)
    scale, shift = torch.empty(24), torch.empty(24)
    out = act(2, 24, 6, 10, wide=8)
    y = F_.conv_fwd(x, wp, 24, 3, 2, 1, scale, shift, relu=True, out=out)
    d, px, pw, ps, psh, py, pst, stream = fake.last("fsb_conv_fwd")
    assert y is out
    assert (d["N"], d["H"], d["W"], d["Cin"], d["Cout"], d["ksize"], d["stride"], d["pad"], d["Ho"], d["Wo"]) == (2, 12, 20, 32, 24, 3, 2, 1, 6, 10)
    assert d["x_cstride"] == 48 and d["y_cstride"] == 32
    assert d["flags"] == _lib.FSB_CONV_RELU | _lib.FSB_CONV_AFFINE
    assert (px, pw, ps, psh, py, pst) == (x.data_ptr(), wp.data_ptr(), scale.data_ptr(), shift.data_ptr(), out.data_ptr(), None)
    # training flavour: fp32 raw output + statistics, shifted origin (FactorizedReduce's second conv)
    # (partial statistic rows: one per CTA of the kernel the library will dispatch; this conv fills columns 24.. of a 48-wide set)
    stats = F_.conv_stats_buffer(x, 24, 1, 2, 0, off=(1, 1), total_C=48)
    q = fake.last("fsb_conv_stats_rows")[0]
    assert (q["Cout"], q["off_h"], q["Ho"]) == (24, 1, 6) and q["flags"] & _lib.FSB_CONV_STATS
    assert tuple(stats.shape) == (3, 96) and stats.dtype == torch.float32
    raw = F_.conv_fwd(x, wp, 24, 1, 2, 0, off=(1, 1), stats=stats, out_f32=True, stats_off=24)
    d, _, _, ps, psh, py, pst, _ = fake.last("fsb_conv_fwd")
    assert raw.dtype == torch.float32 and tuple(raw.shape) == (2, 24, 6, 10)
    assert (d["off_h"], d["off_w"], d["Ho"], d["Wo"]) == (1, 1, 6, 10)
    assert d["flags"] == _lib.FSB_CONV_STATS | _lib.FSB_CONV_OUT_F32 and (ps, psh) == (None, None)
    assert (d["stats_C"], d["stats_off"]) == (48, 24)
    assert (py, pst) == (raw.data_ptr(), stats.data_ptr()) and d["y_cstride"] == raw.stride(3)
    # finalize gets the rows as they are (it adds them in index order itself)
    gamma, beta = torch.empty(48), torch.empty(48)
    F_.bn_finalize(stats, 120, gamma, beta, 1e-5, 0.1, None, None)
    f = fake.last("fsb_bn_finalize")
    assert f[0] == 48 and f[1] == stats.data_ptr() and (f[2], f[3]) == (3, 48) and f[4] == pytest.approx(120.0)


def test_fused_training_unit_marshalling(fake):
    x = act(2, 32, 12, 20)
    wp = torch.empty(2048, dtype=torch.float16)
    gamma, beta, rm, rv = torch.empty(20), torch.empty(20), torch.empty(20), torch.empty(20)
    nbt = torch.zeros((), dtype=torch.int64)
    y, raw, vec, d = F_.conv_bn_act_train_fwd(x, wp, 20, 3, 1, 1, (0, 0), gamma, beta, 1e-5, 0.1, rm, rv, nbt, True)
    a = fake.last("fsb_conv_bn_act_train_fwd")
    desc = a[0]
    assert (desc["N"], desc["Cin"], desc["Cout"], desc["Ho"], desc["Wo"], desc["x_cstride"], desc["y_cstride"]) == (2, 32, 20, 12, 20, 32, 24)
    assert a[1:5] == [x.data_ptr(), wp.data_ptr(), gamma.data_ptr(), beta.data_ptr()]
    assert a[5] == pytest.approx(1e-5) and a[6] == pytest.approx(0.1)
    assert a[7:10] == [rm.data_ptr(), rv.data_ptr(), nbt.data_ptr()]
    assert a[10] == raw.data_ptr() and a[11] == raw.stride(3) == 24 and raw.dtype == torch.float32
    assert a[12] == y.data_ptr() and a[13] == y.stride(3) == 24 and y.dtype == torch.float16
    assert a[14] == vec.data_ptr() and vec.numel() == (6 + 2 * 3) * 20 and a[15] == 1 and a[16] is None and a[17] is None
    assert tuple(y.shape) == tuple(raw.shape) == (2, 20, 12, 20)
    cpu_backend.nhwc_info(y), cpu_backend.nhwc_info(raw, torch.float32)
    # backward
    dy = act(2, 20, 12, 20)
    w = torch.empty(40, 48, 3, 3)[:20, :32]            # active corner of a max-width master weight
    wt = torch.empty(2048, dtype=torch.float16)
    acc = torch.zeros(40, 48, 3, 3)
    dx, dg, db = F_.conv_bn_act_train_bwd(d, x, dy, y, raw, vec, gamma, True, wt, w, True, acc, 1024.0)
    b = fake.last("fsb_conv_bn_act_train_bwd")
    assert b[1:4] == [x.data_ptr(), dy.data_ptr(), dy.stride(3)]
    assert b[4:8] == [y.data_ptr(), y.stride(3), raw.data_ptr(), raw.stride(3)]
    assert b[8:11] == [vec.data_ptr(), gamma.data_ptr(), 1]
    assert b[11:15] == [wt.data_ptr(), w.data_ptr(), w.stride(0), w.stride(1)] and (w.stride(0), w.stride(1)) == (48 * 9, 9)
    assert b[16] == 24 and b[15] is not None and b[17] is not None        # draw (stride = padded Cout) and the 4C scratch vector
    assert tuple(dx.shape) == (2, 32, 12, 20) and b[18] == dx.data_ptr() and b[19] == dx.stride(3)
    assert b[20] == acc.data_ptr() and b[21] == pytest.approx(1024.0) and b[22] is None and b[23] is None
    assert dg.numel() == db.numel() == 20 and dg.data_ptr() != db.data_ptr()
    # [totals 2C | 3 partial rows x 2C | dgamma | dbeta]
    assert dg.data_ptr() - b[17] == (2 + 6) * 20 * 4 and db.data_ptr() - dg.data_ptr() == 20 * 4
    # no dx wanted (first layer): null pointer, no allocation
    dx, _, _ = F_.conv_bn_act_train_bwd(d, x, dy, y, raw, vec, gamma, True, None, w, False, acc, 1024.0)
    b = fake.last("fsb_conv_bn_act_train_bwd")
    assert dx is None and b[18] is None and b[11] is None


def test_bn_backward_marshalling(fake):
    dy, y = act(2, 24, 6, 10), act(2, 24, 6, 10)
    raw = act(2, 24, 6, 10, dtype=torch.float32)
    mean, invstd, gamma = torch.empty(24), torch.empty(24), torch.empty(24)
    draw, dg, db = F_.bn_bwd(dy, y, raw, mean, invstd, gamma, 120, True, 1024.0)
    r = fake.last("fsb_bn_bwd_reduce")
    assert r[0] == 2 * 6 * 10 and r[1] == 24 and r[2] == dy.data_ptr() and r[4] == y.data_ptr() and r[6] == raw.data_ptr()
    assert r[8] == 1 and r[11] == 1      # raw is fp32, relu mask on
    a = fake.last("fsb_bn_bwd_apply")
    assert a[12] == r[12] and a[13] == pytest.approx(120.0)   # same sums buffer (row 0 = totals), count
    assert a[15] == draw.data_ptr() and a[17] == dg.data_ptr() and a[18] == db.data_ptr() and a[20] == 0
    # SyncBN composition: the apply kernel gets the all-reduced buffer and no gamma/beta outputs
    seen = {}

    def allreduce(s):
        seen["in"] = s
        return s
    draw, dg, db = F_.bn_bwd(dy, y, raw, mean, invstd, gamma, 240, True, 1024.0, allreduce=allreduce)
    a = fake.last("fsb_bn_bwd_apply")
    assert a[12] == seen["in"].data_ptr() and a[17] is None and a[18] is None and a[13] == pytest.approx(240.0)
    assert dg.numel() == db.numel() == 24


def test_conv_backward_marshalling(fake):
    dy = act(2, 24, 6, 10)
    w = torch.empty(40, 48, 3, 3)[:24, :32]
    wt = torch.empty(2048, dtype=torch.float16)
    dx = F_.conv_dgrad(dy, w, (2, 32, 12, 20), 32, 24, 3, 2, 1, wpacked_t=wt)
    a = fake.last("fsb_conv_dgrad")
    assert (a[0]["H"], a[0]["W"], a[0]["Ho"], a[0]["Wo"], a[0]["stride"]) == (12, 20, 6, 10, 2)
    assert a[1:3] == [dy.data_ptr(), dy.stride(3)] and a[3] == wt.data_ptr() and a[4:7] == [w.data_ptr(), 48 * 9, 9]
    assert a[7:9] == [dx.data_ptr(), dx.stride(3)] and tuple(dx.shape) == (2, 32, 12, 20)
    x = act(2, 32, 12, 20)
    acc = torch.zeros(40, 48, 3, 3)
    F_.conv_wgrad(x, dy, acc, 32, 24, 3, 2, 1, 1024.0, accumulate_into=acc)
    a = fake.last("fsb_conv_wgrad")
    assert a[1:4] == [x.data_ptr(), dy.data_ptr(), dy.stride(3)] and a[4:8] == [acc.data_ptr(), 48 * 9, 9, 1]
    dw = F_.conv_wgrad(x, dy, acc, 32, 24, 3, 2, 1, 1024.0)
    a = fake.last("fsb_conv_wgrad")
    assert a[7] == 0 and dw.shape == acc.shape and float(dw.abs().sum()) == 0.0   # partial corner: zero-filled full tensor


def test_resize_concat_and_layout_marshalling(fake):
    x = act(1, 32, 8, 16)
    cat = F_.empty_nhwc(1, 48, 16, 32, "cpu")
    F_.bilinear(x, (16, 32), out=cat[:, :32])
    a = fake.last("fsb_bilinear_fwd")
    assert a[:6] == [1, 32, 8, 16, 16, 32] and a[6:10] == [x.data_ptr(), 32, cat.data_ptr(), 48]
    skip = act(1, 16, 16, 32)
    F_.copy_channels(skip, cat[:, 32:])
    a = fake.last("fsb_copy_channels")
    assert a[:2] == [16 * 32, 16] and a[2:6] == [skip.data_ptr(), 16, cat.data_ptr() + 32 * 2, 48]
    logits = act(1, 19, 8, 16)
    out = F_.upsample_logits(logits, (64, 128), dtype=torch.float16)
    a = fake.last("fsb_upsample_logits_nchw")
    assert a[:6] == [1, 19, 8, 16, 64, 128] and a[7] == 24 and a[9] == 0 and out.is_contiguous()
    lab = F_.upsample_argmax(logits, (64, 128))
    assert lab.dtype == torch.uint8 and tuple(lab.shape) == (1, 64, 128)
    img = torch.empty(1, 3, 16, 32)
    xh = F_.to_nhwc_half(img)
    a = fake.last("fsb_nchw_to_nhwc_f16")
    assert a[:4] == [1, 3, 16, 32] and a[5] == 1 and a[7] == 8 and tuple(xh.shape) == (1, 3, 16, 32)
    back = F_.to_nchw(x)
    assert back.dtype == torch.float32 and back.is_contiguous()


def test_weighted_sum_marshalling(fake):
    xs = [act(1, 16, 4, 6) for _ in range(5)]
    wts = torch.empty(5)
    out = F_.wsum_fwd(xs, wts)
    a = fake.last("fsb_wsum_fwd")
    assert a[:3] == [5, 24, 16] and a[5] == wts.data_ptr() and a[6] == out.data_ptr()
    dxs, dw = F_.wsum_bwd(out, xs, wts, [True, False, True, True, True], True, 1024.0)
    assert dxs[1] is None and all(d is not None for i, d in enumerate(dxs) if i != 1) and dw.numel() == 5


def test_weight_packing_and_stem_marshalling(fake):
    w = torch.empty(40, 48, 3, 3)
    packed = F_.pack_conv_weight(w, 32, 24, 3)
    a = fake.last("fsb_pack_conv_weight")
    assert (a[0]["Cin"], a[0]["Cout"], a[0]["ksize"]) == (32, 24, 3) and a[1:4] == [w.data_ptr(), 48 * 9, 9]
    assert packed.dtype == torch.float16 and packed.numel() == 2048 and a[4] == packed.data_ptr()
    img = torch.empty(2, 3, 32, 64)
    ws = torch.empty(32, 3, 3, 3)
    scale, shift = torch.empty(32), torch.empty(32)
    y = F_.stem_conv_nchw(img, ws, scale, shift)
    a = fake.last("fsb_stem_conv_nchw")
    assert a[:4] == [2, 32, 64, 32] and a[4] == img.data_ptr() and a[5] == 1 and a[9] == y.data_ptr() and a[10] == 32
    assert tuple(y.shape) == (2, 32, 16, 32) and a[11] == _lib.FSB_CONV_RELU | _lib.FSB_CONV_AFFINE
    with pytest.raises(ValueError):
        F_.nhwc_info(torch.empty(2, 8, 4, 4, dtype=torch.float16))          # NCHW-contiguous: not NHWC-addressable
    with pytest.raises(C.ArgumentError):
        fake.fsb_conv_fwd(None, torch.empty(1), None, None, None, None, None, None)   # a tensor is not a pointer


def test_fused_criteria_marshalling(fake):
    """N1 wrappers (csrc/loss.cu): shapes, channel strides, label geometry, flags and scalar types as include/fsb200.h declares them"""
    fake.fsb_kth_workspace_bytes = C.CFUNCTYPE(C.c_size_t)(lambda: 4096)
    fake.fsb_loss_rows = C.CFUNCTYPE(C.c_int)(lambda: 7)
    x = act(2, 19, 8, 16)                      # low-resolution logits, channel stride 24
    xt = act(2, 19, 4, 8, wide=8)              # teacher at another resolution inside a wider buffer (stride 32)
    tgt = torch.zeros(2, 64, 128, dtype=torch.int64)
    logp, lse = F_.loss_logp_fwd(x, tgt, (64, 128), 255)
    a = fake.last("fsb_loss_logp_fwd")
    assert a[:6] == [2, 19, 8, 16, 64, 128] and a[6] == x.data_ptr() and a[7] == 24 and a[8] == tgt.data_ptr() and a[9] == 255
    assert a[10] == logp.data_ptr() and a[11] == lse.data_ptr() and logp.dtype == lse.dtype == torch.float32
    assert tuple(logp.shape) == tuple(lse.shape) == (2, 64, 128)
    with pytest.raises(AssertionError):
        F_.loss_logp_fwd(x, tgt.int(), (64, 128), 255)          # labels must be int64, like the reference's target.long()
    k = F_.kth_smallest(logp.reshape(-1), 100)
    a = fake.last("fsb_kth_smallest_f32")
    assert a[0] == logp.data_ptr() and a[1:3] == [2 * 64 * 128, 100] and a[3] == k.data_ptr() and k.dim() == 0
    with pytest.raises(AssertionError):
        F_.kth_smallest(logp.reshape(-1), 0)
    red = F_.ohem_reduce(logp, tgt, 255, 19, thr=k)
    a = fake.last("fsb_ohem_reduce")
    assert a[0] == logp.data_ptr() and a[1] == tgt.data_ptr() and a[2:5] == [2 * 64 * 128, 255, 19] and a[5] == k.data_ptr()
    assert a[7] == red.data_ptr() and red.numel() == 2
    F_.ohem_reduce(logp, tgt, 255, 19)
    assert fake.last("fsb_ohem_reduce")[5] is None              # no threshold: every valid pixel is kept
    coef = torch.empty(())
    dx = F_.loss_ce_bwd(x, tgt, (64, 128), 255, lse, logp, k, coef, 1024.0)
    a = fake.last("fsb_loss_ce_bwd")
    assert a[:6] == [2, 19, 8, 16, 64, 128] and a[7] == 24 and a[9] == 255 and a[10] == lse.data_ptr() and a[11] == logp.data_ptr()
    assert a[12] == k.data_ptr() and a[13] == coef.data_ptr() and a[14] == dx.data_ptr() and a[15] == 24 and a[16] == 1024.0 and a[17] == 0
    F_.loss_ce_bwd(x, tgt, (64, 128), 255, lse, logp, None, coef, 1024.0, out=dx)
    a = fake.last("fsb_loss_ce_bwd")
    assert a[12] is None and a[14] == dx.data_ptr() and a[17] == 1           # accumulate into the given gradient buffer
    total, lse_s, lse_t = F_.loss_kl_fwd(x, xt, (64, 128))
    a = fake.last("fsb_loss_kl_fwd")
    assert a[:8] == [2, 19, 8, 16, 4, 8, 64, 128] and a[8] == x.data_ptr() and a[9] == 24 and a[10] == xt.data_ptr() and a[11] == 32
    assert a[12] == lse_s.data_ptr() and a[13] == lse_t.data_ptr() and total.dim() == 0
    dxs = F_.loss_kl_bwd(x, xt, (64, 128), lse_s, lse_t, coef, 1024.0)
    a = fake.last("fsb_loss_kl_bwd")
    assert a[:8] == [2, 19, 8, 16, 4, 8, 64, 128] and a[9] == 24 and a[11] == 32 and a[14] == coef.data_ptr() and a[15] == dxs.data_ptr()
    assert a[16] == 24 and a[17] == 1024.0 and a[18] == 0 and tuple(dxs.shape) == (2, 19, 8, 16)


def test_flat_step_tail_marshalling(fake):
    """csrc/optim.cu through functional.flat_* with the tables optim.FlatTables builds for a (stand-in) flat gradient buffer"""
    import numpy as np
    from fasterseg_b200 import optim as FO
    fake.fsb_flat_chunk = C.CFUNCTYPE(C.c_int)(lambda: 4096)

    class Flat:      # the part of graphed.FlatGrads the tables need
        def __init__(self, params):
            self.params, self.offsets, total = params, {}, 0
            for p in params:
                self.offsets[id(p)] = total
                total += (p.numel() + 3) // 4 * 4
            self.G = torch.zeros(total)

    params = [torch.nn.Parameter(torch.zeros(n)) for n in (5, 4096, 4097, 10000)]
    flat = Flat(params)
    t = FO.FlatTables(flat)
    assert t.nblocks == 1 + 1 + 2 + 3 and t.map.dtype == torch.int32 and tuple(t.map.shape) == (7, 2)
    assert t.map.tolist() == [[0, 0], [1, 0], [2, 0], [2, 1], [3, 0], [3, 1], [3, 2]]
    seg = np.frombuffer(t.segs.numpy().tobytes(), dtype=[("p", "<u8"), ("off", "<u4"), ("n", "<u4")])
    assert seg["p"].tolist() == [p.data_ptr() for p in params] and seg["n"].tolist() == [5, 4096, 4097, 10000]
    assert seg["off"].tolist() == [0, 8, 8 + 4096, 8 + 4096 + 4100] and t.segs.numel() == 4 * 16        # struct FlatSeg is 16 bytes
    assert t.pointers_valid()
    t.set_live(np.array([1, 0, 1, 1], dtype=np.uint8))
    assert t.live.tolist() == [1, 0, 1, 1]
    extra = torch.empty(1)
    F_.flat_grad_norm(t.map, t.nblocks, t.segs, t.live, flat.G, t.partial, extra, 5, t.norm)
    a = fake.last("fsb_flat_grad_norm")
    assert a[0] == t.map.data_ptr() and a[1] == 7 and a[2] == t.segs.data_ptr() and a[3] == t.live.data_ptr() and a[4] == flat.G.data_ptr()
    assert a[5] == t.partial.data_ptr() and a[6] == extra.data_ptr() and a[7] == 5.0 and a[8] == t.norm.data_ptr() and t.partial.numel() == 7
    F_.flat_grad_norm(t.map, t.nblocks, t.segs, t.live, flat.G, t.partial, None, 5.0, t.norm)
    assert fake.last("fsb_flat_grad_norm")[6] is None
    F_.flat_scale(t.map, t.nblocks, t.segs, t.live, flat.G, t.norm[1:])
    a = fake.last("fsb_flat_scale")
    assert a[1] == 7 and a[4] == flat.G.data_ptr() and a[5] == t.norm.data_ptr() + 4            # the clip coefficient is norm[1]
    M = torch.zeros_like(flat.G)
    F_.flat_sgd(t.map, t.nblocks, t.segs, t.live, flat.G, M, 0.05, 0.9, 5e-4)
    a = fake.last("fsb_flat_sgd")
    assert a[4] == flat.G.data_ptr() and a[5] == M.data_ptr() and a[6] == pytest.approx(0.05) and a[7] == pytest.approx(0.9)
    assert a[8] == pytest.approx(5e-4)
    params[2].data = torch.zeros(4097)               # storage re-allocated behind the table's back: every pointer is compared
    assert not t.pointers_valid()
    t2 = FO._tables(flat)                            # ... and the tables are rebuilt for the new storage
    assert t2 is not t and t2.pointers_valid() and int(t2.ptrs[2]) == params[2].data_ptr()
