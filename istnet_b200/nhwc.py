"""Channels-last building blocks on the B200 kernels: the "conv -> BatchNorm -> activation" unit with a hand-written
backward (tcgen05 GEMMs for forward / dgrad / wgrad, fused HBM passes in between).

Everything here works on FP32 tensors [B,H,W,C] (or [1,1,rows,C] for point sets) and on their bf16 (hi, lo) operand
pairs; see include/istnet_b200.h §3-§4 for the kernels.  The classes mirror what PyTorch autograd records for the
reference modules (cuDNN conv fwd/dgrad/wgrad, BN fwd/bwd, ReLU/PReLU, Dropout2d), but as an explicit tape.
"""
import ctypes
import os
import weakref

import torch

from . import _C
from ._C import c_float, c_int, c_ll, c_void_p, ptr

NULL = c_void_p(0)
# bf16 operand planes per FP32 tensor: 3 -> six tensor-core products, FP32-level accuracy (default, needed for the 1e-4
# parity bar in train mode); 2 -> three products, ~3e-6 per product, twice the tensor-core throughput.
NSPLIT = int(os.environ.get("ISTNET_NSPLIT", "3"))
# The backward contractions (dgrad / wgrad) use the first NSPLIT_BWD planes only: 3 products, ~3e-6 per product —
# orders of magnitude below the FP32 noise floor of the gradients themselves (tests/golden `referr_grad_*`).
NSPLIT_BWD = min(NSPLIT, int(os.environ.get("ISTNET_NSPLIT_BWD", "2")))


def _ptrs(tensors):
    """C array of device pointers (NULL for None) for the `const float *const *` arguments of csrc/heads.cu."""
    return (ctypes.c_void_p * len(tensors))(*[t.data_ptr() if t is not None else None for t in tensors])


def _p(t):
    return ptr(t) if t is not None else NULL


def pad8(c):
    return (c + 7) // 8 * 8


class Act:
    """An activation tensor in channels-last layout: optional FP32 copy + optional bf16 operand planes [NSPLIT,B,H,W,cs]."""

    __slots__ = ("f32", "pl", "B", "H", "W", "C")

    def __init__(self, B, H, W, C, f32=None, pl=None):
        self.B, self.H, self.W, self.C, self.f32, self.pl = B, H, W, C, f32, pl

    @property
    def P(self):
        return self.B * self.H * self.W

    @property
    def cs(self):
        return self.pl.shape[-1]

    @property
    def hi(self):
        return self.pl[0] if self.pl is not None else None


def empty_planes(B, H, W, C, dev, cs=None, nsplit=None):
    return torch.empty(nsplit or NSPLIT, B, H, W, cs or pad8(C), dtype=torch.bfloat16, device=dev)


def _pl_args(pl):
    """(pointer, plane stride in elements, nsplit) of an operand-plane tensor (or nulls)."""
    if pl is None:
        return NULL, c_ll(0), c_int(NSPLIT)
    return ptr(pl), c_ll(pl.stride(0)), c_int(pl.shape[0])


# ----------------------------------------------------------------------------------------- thin kernel wrappers
def split(x_f32, P, C, pl, ch_off=0, HW=1, nchw=False):
    _C.call("split", ptr(x_f32), c_ll(P), c_int(C), c_ll(HW), c_int(1 if nchw else 0), *_pl_args(pl), c_int(pl.shape[-1]), c_int(ch_off))


def prep_weight(w, transpose=False, nsplit=None, im2col=False):
    """Conv / linear weight [co, ci, kh, kw] (or [co, ci(,1)]) -> bf16 operand planes, one fused kernel:
    transpose=False: forward operand [ns, tap, co, pad8(ci)];  transpose=True: data-gradient operand [ns, flipped tap, ci, pad8(co)];
    im2col=True (strided convs run as a 1x1 GEMM over patches): [ns, 1, co, pad8(kh*kw*ci)] / transposed [ns, 1, kh*kw*ci, pad8(co)]."""
    if w.dim() == 2:
        co, ci, kh, kw = w.shape[0], w.shape[1], 1, 1
    elif w.dim() == 3:
        co, ci, kh, kw = w.shape[0], w.shape[1], w.shape[2], 1
    else:
        co, ci, kh, kw = w.shape
    taps = kh * kw
    if im2col:
        shape = (1, taps * ci, pad8(co)) if transpose else (1, co, pad8(taps * ci))
    else:
        shape = (taps, ci, pad8(co)) if transpose else (taps, co, pad8(ci))
    if WEIGHT_BANK and not transpose and isinstance(w, torch.nn.Parameter) and w.is_cuda and weight_bank(w.device) is not None:
        st = (1, taps * ci, pad8(co)) if im2col else (taps, ci, pad8(co))
        return weight_bank(w.device).get(w, nsplit or NSPLIT, im2col, False, (co, ci, kh, kw), (shape, st))[0]
    pl = torch.empty(nsplit or NSPLIT, *shape, dtype=torch.bfloat16, device=w.device)
    _C.call("prep_weight", ptr(w), c_int(co), c_int(ci), c_int(kh), c_int(kw), c_int(1 if transpose else 0), c_int(1 if im2col else 0),
            *_pl_args(pl), c_int(shape[-1]))
    return pl


def prep_weight_pair(w, im2col=False, nsplit=None):
    """Forward operand (nsplit planes) and data-gradient operand (NSPLIT_BWD planes, transposed / flipped) of one weight, one launch."""
    if w.dim() == 2:
        co, ci, kh, kw = w.shape[0], w.shape[1], 1, 1
    elif w.dim() == 3:
        co, ci, kh, kw = w.shape[0], w.shape[1], w.shape[2], 1
    else:
        co, ci, kh, kw = w.shape
    taps = kh * kw
    sf = (1, co, pad8(taps * ci)) if im2col else (taps, co, pad8(ci))
    st = (1, taps * ci, pad8(co)) if im2col else (taps, ci, pad8(co))
    if WEIGHT_BANK and isinstance(w, torch.nn.Parameter) and w.is_cuda and weight_bank(w.device) is not None:
        return weight_bank(w.device).get(w, nsplit or NSPLIT, im2col, True, (co, ci, kh, kw), (sf, st))
    pf = torch.empty(nsplit or NSPLIT, *sf, dtype=torch.bfloat16, device=w.device)
    pt = torch.empty(min(nsplit or NSPLIT, NSPLIT_BWD), *st, dtype=torch.bfloat16, device=w.device)
    _C.call("prep_weight_pair", ptr(w), c_int(co), c_int(ci), c_int(kh), c_int(kw), c_int(1 if im2col else 0), *_pl_args(pf), c_int(sf[-1]),
            *_pl_args(pt), c_int(st[-1]))
    return pf, pt


# ---- all weights of a model re-laid in ONE launch per step
# A convolution weight changes once per optimizer step, yet round 1 re-laid it right in front of every GEMM (~130 launches of 3-9 us
# on the dependent chain of each stream).  The bank remembers every nn.Parameter that went through prep_weight_pair / prep_weight
# (persistent destination planes, one table entry each); `begin_step()` — called at the start of IST_Net / PoseNetGT.forward —
# refreshes ALL of them with one launch of istnet_prep_weight_batch, after which the per-layer calls are table look-ups.
# Temporaries (weight slices, the head's correction matrix) are not nn.Parameters and keep the immediate path.
class _PrepEntry(ctypes.Structure):
    _fields_ = [("w", c_void_p), ("pf", c_void_p), ("pt", c_void_p), ("stride_f", ctypes.c_longlong), ("stride_t", ctypes.c_longlong),
                ("co", ctypes.c_int), ("ci", ctypes.c_int), ("kh", ctypes.c_int), ("kw", ctypes.c_int), ("im2col", ctypes.c_int),
                ("ns_f", ctypes.c_int), ("cs_f", ctypes.c_int), ("ns_t", ctypes.c_int), ("cs_t", ctypes.c_int), ("block0", ctypes.c_int)]


PREP_CHUNK = 2048  # ISTNET_PREP_CHUNK
WEIGHT_BANK = os.environ.get("ISTNET_WEIGHT_BANK", "1") != "0"


class WeightBank:
    def __init__(self):
        self.entries = {}   # (id(param), nsplit, im2col) -> dict(w=weakref, pf, pt, dims, im2col, fresh, ver)
        self.table = None   # device copy of the entry table
        self.ptrs = None    # data pointers the table was built with
        self.blocks = self.n = 0
        self.step_id = 0    # bumped by begin_step / invalidate; entries refreshed in the current step carry the same id
        self.dirty = True
        self._tables = []

    def _live(self):
        dead = [k for k, e in self.entries.items() if e["w"]() is None]
        for k in dead:
            del self.entries[k]
            self.dirty = True
        return [(e, e["w"]()) for e in self.entries.values()]

    def _build(self, dev, live):
        arr = (_PrepEntry * len(live))()
        b0 = 0
        for a, (e, w) in zip(arr, live):
            co, ci, kh, kw = e["dims"]
            a.w, a.pf, a.pt = w.data_ptr(), e["pf"].data_ptr(), (e["pt"].data_ptr() if e["pt"] is not None else None)
            a.stride_f, a.stride_t = e["pf"].stride(0), (e["pt"].stride(0) if e["pt"] is not None else 0)
            a.co, a.ci, a.kh, a.kw, a.im2col = co, ci, kh, kw, int(e["im2col"])
            a.ns_f, a.cs_f = e["pf"].shape[0], e["pf"].shape[-1]
            a.ns_t, a.cs_t = (e["pt"].shape[0], e["pt"].shape[-1]) if e["pt"] is not None else (0, 0)
            a.block0 = b0
            b0 += (co * ci * kh * kw + PREP_CHUNK - 1) // PREP_CHUNK
        self.table = torch.frombuffer(bytearray(bytes(arr)), dtype=torch.uint8).to(dev)
        self._tables.append(self.table)  # a CUDA graph captured earlier keeps launching with ITS table: never freed while the bank lives
        self.blocks, self.n = b0, len(live)
        self.ptrs = tuple(w.data_ptr() for _, w in live)
        self.dirty = False

    def invalidate(self):
        """The weights changed outside autograd's version counting (FlatAdam's kernel): nothing refreshed so far is valid."""
        self.step_id += 1

    def begin_step(self, dev):
        """Re-lays every registered weight (one launch).  No-op until a first step has registered the model's weights."""
        self.step_id += 1
        if not WEIGHT_BANK or not self.entries:
            return
        live = self._live()
        if not live:
            return
        settled = (not self.dirty) and self.ptrs == tuple(w.data_ptr() for _, w in live)
        if not settled:
            if torch.cuda.is_current_stream_capturing():
                return  # table not settled (capture without warm-up): this step uses the immediate path
            self._build(dev, live)  # new entries, or the optimizer re-pointed the parameters into flat buffers
        _C.call("prep_weight_batch", ptr(self.table), c_int(self.n), c_int(self.blocks))
        for e, w in live:
            e["fresh"], e["ver"] = self.step_id, w._version

    def get(self, w, nsplit, im2col, pair, dims, shapes):
        """Planes of parameter `w` for this step: from the bank if begin_step refreshed them (and the parameter has not been
        modified since), else immediate prep + registration."""
        key = (id(w), nsplit, bool(im2col))
        e = self.entries.get(key)
        if e is not None and e["w"]() is w and (e["pt"] is not None or not pair):
            if e.get("fresh") == self.step_id and e.get("ver") == w._version and not self.dirty:
                return e["pf"], e["pt"]
            pf, pt = e["pf"], e["pt"]
        else:
            sf, st = shapes
            pf = torch.empty(nsplit, *sf, dtype=torch.bfloat16, device=w.device)
            pt = torch.empty(min(nsplit, NSPLIT_BWD), *st, dtype=torch.bfloat16, device=w.device) if pair else None
            if not torch.cuda.is_current_stream_capturing():  # persistent buffers must not come from a graph's private pool
                self.entries[key] = {"w": weakref.ref(w), "pf": pf, "pt": pt, "dims": dims, "im2col": im2col}
                self.dirty = True
        co, ci, kh, kw = dims
        if pt is not None:
            _C.call("prep_weight_pair", ptr(w), c_int(co), c_int(ci), c_int(kh), c_int(kw), c_int(1 if im2col else 0), *_pl_args(pf), c_int(pf.shape[-1]),
                    *_pl_args(pt), c_int(pt.shape[-1]))
        else:
            _C.call("prep_weight", ptr(w), c_int(co), c_int(ci), c_int(kh), c_int(kw), c_int(0), c_int(1 if im2col else 0), *_pl_args(pf), c_int(pf.shape[-1]))
        return pf, pt


_CURRENT_BANK = {}  # device index -> bank of the model whose forward ran last on that device


def weight_bank(dev):
    return _CURRENT_BANK.get(dev.index if dev.index is not None else torch.cuda.current_device())


def begin_step(model, dev):
    """Call once at the start of a model forward (IST_Net / PoseNetGT): refreshes the operand planes of all registered weights of
    THIS model with one launch.  The bank lives on the model object, so its buffers and table die with the model (and a CUDA
    graph captured for the model only ever references memory the model owns)."""
    if dev.type != "cuda" or not WEIGHT_BANK:
        return
    bank = model.__dict__.get("_istnet_weight_bank")
    if bank is None:
        bank = WeightBank()
        object.__setattr__(model, "_istnet_weight_bank", bank)
    _CURRENT_BANK[dev.index if dev.index is not None else torch.cuda.current_device()] = bank
    bank.begin_step(dev)


def invalidate_weights(dev):
    """An optimizer that writes parameters through raw pointers (parallel.FlatAdam) calls this after its step."""
    if dev.type == "cuda":
        b = weight_bank(dev)
        if b is not None:
            b.invalidate()


def pick_box(H, W):
    if H == 1:
        return 128, 1
    if W % 16 == 0 and H % 8 == 0:
        return 16, 8
    return 8, 8


FUSE_BN_STATS = os.environ.get("ISTNET_FUSE_BN_STATS", "1") != "0"
# ReLU backward + bias gradient + operand split of a bias+ReLU layer inside the epilogue of the data-gradient GEMM above it
FUSE_RELU_BWD = os.environ.get("ISTNET_FUSE_RELU_BWD", "1") != "0"
# conv+BN+ReLU layer below: ReLU mask and the BatchNorm-backward reduction (sum g, sum g*y) in that epilogue; only the apply pass remains
FUSE_BN_BWD = os.environ.get("ISTNET_FUSE_BN_BWD", "0") != "0"  # measured: no gain (1133 vs 1136 inst/s), see DESIGN.md section 9
WGRAD_SIDE_STREAM = os.environ.get("ISTNET_WGRAD_STREAM", "1") != "0"
_SIDE = {}
_PENDING_JOINS = []


def _side_stream(dev):
    key = (dev, torch.cuda.current_stream(dev).cuda_stream)
    st = _SIDE.get(key)
    if st is None:
        st = _SIDE[key] = torch.cuda.Stream(dev)
    return st


def side_stream_for(dev, slot):
    """A per-(current stream, slot) side stream (slot 0 is used by the weight-gradient overlap)."""
    if dev.type != "cuda":
        return None
    key = (dev, torch.cuda.current_stream(dev).cuda_stream, slot)
    st = _SIDE.get(key)
    if st is None:
        st = _SIDE[key] = torch.cuda.Stream(dev)
    return st


def join_side_streams():
    """Makes every stream that forked a weight-gradient side stream wait for it (called once per backward of a tape)."""
    while _PENDING_JOINS:
        main, side = _PENDING_JOINS.pop()
        main.wait_stream(side)


# bench.py sets PROFILE = [] to time every tensor-core launch with CUDA events on the launching stream.  Each profiled
# call is re-issued PROFILE_REPS times back to back between the two events (the kernels are pure functions of their
# inputs), so that the measurement is the device-side duration and not the host's launch latency on an idle stream.
# entries: (kernel name, start event, end event, repetitions, algorithmic FLOPs, nsplit, shape description)
PROFILE = None
PROFILE_REPS = 3


def _timed(name, flops, nsplit, launch, desc=""):
    launch()
    _timed_only(name, flops, nsplit, launch, desc)


def _scratch_fin(fin):
    """Copy of a BatchNorm-statistics descriptor whose outputs go to scratch (no running-statistics / counter update)."""
    rep = _C.Fin.from_buffer_copy(fin)
    if fin.kind == _C.FIN_BN_STATS:
        rep.running_mean = rep.running_var = rep.num_batches_tracked = None
    return rep


def _timed_only(name, flops, nsplit, launch, desc=""):
    if PROFILE is None:
        return
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(PROFILE_REPS):
        launch()
    e1.record()
    PROFILE.append((name, e0, e1, PROFILE_REPS, flops, nsplit, desc))


def conv_gemm(x, w_pl, cout, kh, kw, bias=None, relu=False, out_f32=None, out_pl=None, stat_part=None, mask_hi=None, stat_y=None, fin=None,
              bias_group=0):
    """x: Act with operand planes; writes out_f32 [B,H,W,cout] and/or out_pl.  With stat_part (float buffer of
    >= 2*FIN_ROWS*cout) the epilogue also leaves per-CTA BN-statistics partials there; with `fin` (a _C.Fin) the kernel's last
    CTA finishes that reduction itself (BatchNorm statistics / bias gradient), no finalize launch.  Returns the CTA count G."""
    bw, bh = pick_box(x.H, x.W)
    ns = min(x.pl.shape[0], w_pl.shape[0])  # planes are nested: the first n planes of an operand are its n-plane representation
    grid = ctypes.c_int(0)
    args = (
        ptr(x.pl), c_ll(x.pl.stride(0)), c_int(x.B), c_int(x.H), c_int(x.W), c_int(x.C), c_int(x.cs), ptr(w_pl),
        c_ll(w_pl.stride(0)), c_int(cout), c_int(w_pl.shape[-1]), c_int(kh), c_int(kw), c_int(ns), _p(bias), c_int(1 if relu else 0),
        _p(out_f32), c_int(out_f32.shape[-1] if out_f32 is not None else 0), *_pl_args(out_pl), c_int(out_pl.shape[-1] if out_pl is not None else 0),
        c_int(bw), c_int(bh), _p(stat_part), ctypes.byref(grid), _p(mask_hi), c_int(mask_hi.shape[-1] if mask_hi is not None else 0),
        _p(stat_y), c_int(stat_y.shape[-1] if stat_y is not None else 0),
    )
    fin_ref = ctypes.byref(fin) if fin is not None else NULL
    _C.call("conv_gemm", *args, fin_ref, c_int(bias_group))
    if PROFILE is not None:  # timed repeats must not update running statistics again: same kernel, statistics into scratch outputs
        rep = _scratch_fin(fin) if fin is not None else None
        rep_ref = ctypes.byref(rep) if rep is not None else NULL
        _timed_only("conv_gemm_tc_kernel", 2.0 * x.P * cout * x.C * kh * kw, ns, lambda: _C.call("conv_gemm", *args, rep_ref, c_int(bias_group)),
                    f"P={x.P} {x.H}x{x.W} cin={x.C} cout={cout} k={kh}")
    return grid.value


def conv_wgrad(dy_pl, cout, x, kh, kw):
    """grad_w [cout, cin, kh, kw] from dy operand planes [NSPLIT,B,H,W,cs] and x: Act (planes)."""
    dev = x.pl.device
    xpl = x.pl[: dy_pl.shape[0]]  # planes are nested: the first n planes are the n-plane representation
    ks = _C.lib().istnet_wgrad_ksplit(x.B, x.H, x.W, cout, x.C, kh, kw, xpl.shape[0])
    ws = torch.empty(ks * kh * kw * cout * x.C, dtype=torch.float32, device=dev)
    gw = torch.empty(cout, x.C, kh, kw, dtype=torch.float32, device=dev)
    bw, bh = (64, 1) if x.H == 1 else (8, 8)
    args = (
        ptr(dy_pl), c_ll(dy_pl.stride(0)), c_int(dy_pl.shape[-1]), ptr(xpl), c_ll(xpl.stride(0)), c_int(x.cs), c_int(xpl.shape[0]),
        c_int(x.B), c_int(x.H), c_int(x.W), c_int(cout), c_int(x.C), c_int(kh), c_int(kw), ptr(ws), c_int(ks), ptr(gw), c_int(bw), c_int(bh),
    )
    _timed("wgrad_tc_kernel", 2.0 * x.P * cout * x.C * kh * kw, xpl.shape[0], lambda: _C.call("conv_wgrad", *args),
           f"P={x.P} {x.H}x{x.W} cin={x.C} cout={cout} k={kh} ks={ks}")
    return gw


class BnState:
    """Per-call BatchNorm quantities: batch (train) or running (eval) mean / invstd + affine parameters."""

    __slots__ = ("mean", "invstd", "gamma", "beta", "batch")

    def __init__(self, mean, invstd, gamma, beta, batch=True):
        self.mean, self.invstd, self.gamma, self.beta, self.batch = mean, invstd, gamma, beta, batch


def bn_uses_batch_stats(bn, training):
    return bn is not None and training and bn.training  # each BatchNorm module's own flag decides, as in nn.BatchNorm2d.forward


# ---- BatchNorm momentum as DEVICE scalars.  The reference rewrites `bn.momentum` on every BatchNorm module each iteration
# (BNMomentumScheduler, utils/scheduler.py:277-303, utils/solver.py:48-49,91-92).  A kernel argument passed by value would be
# frozen into a captured CUDA graph, so the statistics kernels read momentum from a per-device table instead: every
# nn.BatchNorm2d gets a slot, `momentum_ptr` keeps the slot equal to the module's current Python value (eager path) and
# `refresh_momentum` does the same for a set of modules with one host->device copy before a graph replay.
# momentum=None (cumulative moving average) is stored as -1 and resolved on the device from num_batches_tracked.
_MOM_SLOTS = 4096
_mom_tables = {}


class _MomTable:
    def __init__(self, dev):
        self.dev = torch.full((_MOM_SLOTS,), 0.1, dtype=torch.float32, device=dev)
        self.host = torch.full((_MOM_SLOTS,), 0.1, dtype=torch.float32).pin_memory()
        self.mirror = [None] * _MOM_SLOTS  # Python copy of what the device holds (no tensor op on the per-call path)
        self.used = 0
        self.copied = None  # event after the last host->device copy of the table: `host` is rewritten only once that copy is done


def _mom_value(bn):
    return -1.0 if bn.momentum is None else float(bn.momentum)


def _mom_slot(bn, dev):
    key = dev.index if dev.index is not None else torch.cuda.current_device()
    tab = _mom_tables.get(key)
    if tab is None:
        tab = _mom_tables[key] = _MomTable(dev)
    slot = getattr(bn, "_istnet_mom_slot", None)
    if slot is None or slot[0] != key:
        if tab.used >= _MOM_SLOTS:
            raise RuntimeError("istnet_b200: more than %d BatchNorm modules on one device" % _MOM_SLOTS)
        slot = (key, tab.used)
        tab.used += 1
        object.__setattr__(bn, "_istnet_mom_slot", slot)
    return tab, slot[1]


def momentum_ptr(bn, dev):
    """Device address of this module's momentum scalar, brought up to date with `bn.momentum`."""
    tab, i = _mom_slot(bn, dev)
    v = _mom_value(bn)
    if tab.mirror[i] != v:
        if torch.cuda.is_current_stream_capturing():
            raise RuntimeError("istnet_b200: a BatchNorm momentum changed while a CUDA graph was being captured; "
                               "run one eager step (or refresh_momentum) with the new value first")
        tab.mirror[i] = v
        tab.host[i] = v
        tab.dev[i : i + 1].fill_(v)
    return c_void_p(tab.dev.data_ptr() + 4 * i)


def momentum_tensor(bn, dev):
    """The same scalar as a 0-dim tensor view (for the few statistics updates still written in torch)."""
    momentum_ptr(bn, dev)
    tab, i = _mom_slot(bn, dev)
    return tab.dev[i]


def refresh_momentum(bn_modules, dev):
    """Makes the device scalars of `bn_modules` equal to their current `.momentum` (one pinned host->device copy if any
    changed).  GraphedTrainStep calls this before every replay, so a scheduler update is honoured by the captured step."""
    tab, dirty = None, False
    for bn in bn_modules:
        tab, i = _mom_slot(bn, dev)
        v = _mom_value(bn)
        if tab.mirror[i] != v:
            if not dirty and tab.copied is not None:
                tab.copied.synchronize()  # a loop running ahead of the device: the previous table copy must have read `host`
            tab.mirror[i] = v
            tab.host[i] = v
            dirty = True
    if dirty:
        tab.dev.copy_(tab.host, non_blocking=True)
        tab.copied = tab.copied or torch.cuda.Event()
        tab.copied.record()
    return dirty


def bn_begin(bn, C, training, P, dev):
    """State of one BatchNorm call: reads eps / running statistics / training from the nn.BatchNorm2d at call time.
    Train mode: returns (BnState with fresh mean / invstd tensors, _C.Fin) — the descriptor goes to the kernel that produces
    the statistics partials (conv_gemm epilogue, sa_gather_l0, bn_stats), whose last CTA fills mean / invstd and updates the
    running statistics (momentum from the device table) and num_batches_tracked.  Eval mode: (running-statistics state, None)."""
    if not bn_uses_batch_stats(bn, training):
        return BnState(bn.running_mean, torch.rsqrt(bn.running_var + bn.eps), bn.weight, bn.bias, batch=False), None
    mean = torch.empty(C, dtype=torch.float32, device=dev)
    invstd = torch.empty(C, dtype=torch.float32, device=dev)
    track = bn.track_running_stats and bn.running_mean is not None
    nbt = bn.num_batches_tracked if (track and bn.num_batches_tracked is not None) else None
    fin = _C.Fin()
    fin.kind, fin.tickets, fin.P, fin.eps = _C.FIN_BN_STATS, _C.tickets(dev).value, P, bn.eps
    fin.mean, fin.invstd = mean.data_ptr(), invstd.data_ptr()
    if track:
        fin.momentum = momentum_ptr(bn, dev).value
        fin.running_mean, fin.running_var = bn.running_mean.data_ptr(), bn.running_var.data_ptr()
        if nbt is not None and nbt.is_cuda:
            fin.num_batches_tracked = nbt.data_ptr()
        elif nbt is not None:
            if bn.momentum is None:
                raise RuntimeError("istnet_b200: momentum=None needs num_batches_tracked on the device")
            bn.num_batches_tracked += 1
    return BnState(mean, invstd, bn.weight, bn.bias, batch=True), fin


def stat_scratch(C, dev, nacc=2):
    """Partial-sum scratch of a statistics epilogue (both ticket levels, csrc/ticket.cuh)."""
    return torch.empty(nacc * _C.FIN_ROWS * C, dtype=torch.float32, device=dev)


def bn_stats(y, P, C, fin):
    """Train-mode statistics of an existing FP32 tensor (one launch; used when the producing GEMM did not carry them)."""
    _C.call("bn_stats_fin", ptr(y), c_ll(P), c_int(C), ptr(stat_scratch(C, y.device)), ctypes.byref(fin))


def bn_act_split(y, P, C, HW, bn=None, res=None, res_bn=None, act=0, prelu=None, noise=None, out_f32=None, out_pl=None, ch_off=0):
    _C.call(
        "bn_act_split", ptr(y), c_ll(P), c_int(C), c_ll(HW), _p(bn.mean if bn else None), _p(bn.invstd if bn else None),
        _p(bn.gamma if bn else None), _p(bn.beta if bn else None), _p(res), _p(res_bn.mean if res_bn else None),
        _p(res_bn.invstd if res_bn else None), _p(res_bn.gamma if res_bn else None), _p(res_bn.beta if res_bn else None), c_int(act),
        _p(prelu), _p(noise), _p(out_f32), *_pl_args(out_pl), c_int(out_pl.shape[-1] if out_pl is not None else 0), c_int(ch_off),
    )


def bn_act_bwd(dz, dz2, y, P, C, HW, bn, act, prelu, z_hi, noise, dy_pl=None, dy_f32=None, g_out=None, argmax=None, ns=0):
    """Returns (ws, sum_g, sum_gx): ws = 3*C doubles [sum g | sum g*xhat | PReLU slope partials]; sum_g / sum_gx are FP32 copies of
    the first two thirds in tensors of their own (the BatchNorm bias / weight gradients: handed to autograd as they are)."""
    ws = torch.empty(3 * C, dtype=torch.float64, device=dz.device)
    sg = torch.empty(C, dtype=torch.float32, device=dz.device)
    sgx = torch.empty(C, dtype=torch.float32, device=dz.device)
    part = torch.empty(_C.lib().istnet_reduce_ws_floats(c_ll(P), C, 3), dtype=torch.float32, device=dz.device)
    _C.call(
        "bn_act_bwd", ptr(dz), _p(dz2), _p(y), c_ll(P), c_int(C), c_ll(HW), _p(bn.mean if bn else None), _p(bn.invstd if bn else None),
        _p(bn.gamma if bn else None), _p(bn.beta if bn else None), c_int(act), _p(prelu), _p(z_hi), c_int(z_hi.shape[-1] if z_hi is not None else 0),
        _p(noise), c_int(1 if (bn is not None and bn.batch) else 0), _p(argmax), c_int(ns), ptr(part), ptr(ws), *_pl_args(dy_pl), c_int(dy_pl.shape[-1] if dy_pl is not None else 0), _p(dy_f32), _p(g_out), ptr(sg), ptr(sgx),
        _C.tickets(dz.device),
    )
    return ws, sg, sgx


def upsample2x(x_f32, B, H, W, C, out_pl):
    _C.call("upsample2x_split", ptr(x_f32), c_int(B), c_int(H), c_int(W), c_int(C), *_pl_args(out_pl), c_int(out_pl.shape[-1]), NULL)


def upsample2x_bwd(dout, B, H, W, C):
    dx = torch.empty(B, H, W, C, dtype=torch.float32, device=dout.device)
    _C.call("upsample2x_bwd", ptr(dout), c_int(B), c_int(H), c_int(W), c_int(C), ptr(dx))
    return dx


def im2col(x_f32, nchw, B, H, W, C, k, stride, pad):
    Ho, Wo = (H + 2 * pad - k) // stride + 1, (W + 2 * pad - k) // stride + 1
    K = k * k * C
    pl = empty_planes(B, Ho, Wo, K, x_f32.device)
    _C.call("im2col_split", ptr(x_f32), c_int(1 if nchw else 0), c_int(B), c_int(H), c_int(W), c_int(C), c_int(k), c_int(k), c_int(stride),
            c_int(pad), *_pl_args(pl), c_int(pl.shape[-1]))
    return Act(B, Ho, Wo, K, None, pl)


def col2im(dcol, B, H, W, C, k, stride, pad, dx=None):
    acc = dx is not None
    if dx is None:
        dx = torch.empty(B, H, W, C, dtype=torch.float32, device=dcol.device)
    _C.call("col2im", ptr(dcol), c_int(B), c_int(H), c_int(W), c_int(C), c_int(k), c_int(k), c_int(stride), c_int(pad), ptr(dx), c_int(1 if acc else 0))
    return dx


# ----------------------------------------------------------------------------------------- the conv+BN+act unit
ACT_NONE, ACT_RELU, ACT_PRELU, ACT_RELU_MAXROWS = 0, 1, 2, 3


class ConvUnit:
    """One `conv (+bias) -> [BatchNorm] -> [residual add] -> activation -> [Dropout2d scale]` step with its tape.

    stride-1 "same" convolutions run as implicit GEMM on the activation pair; strided ones (conv1, layer2.0) go through
    im2col (the patch matrix is then the GEMM's A operand and the wgrad's B operand)."""

    def __init__(self, conv_w, conv_b, bn, act, prelu=None, k=1, stride=1, pad=None, nsplit=None, nsplit_out=None):
        self.w, self.b, self.bn, self.act, self.prelu = conv_w, conv_b, bn, act, prelu
        self.ns = min(nsplit or NSPLIT, NSPLIT)          # operand planes of this unit's forward contraction (input and weight)
        self.ns_out = min(nsplit_out or NSPLIT, NSPLIT)  # planes written for the consumer of this unit's output
        self.k, self.stride = k, stride
        self.pad = k // 2 if pad is None else pad
        self.cout = conv_w.shape[0]
        # rows per instance when this 1x1 unit's input is [f | mean_over_the_instance(f).expand] (the estimators' global feature,
        # ist_net.py:172-173,257-258,325-326) and only f is passed in: conv([f | m]) = W_a f + (W_b m + b), the bracket is a per-instance bias
        self.glob = None

    # ---- forward
    def forward(self, x, training, record, noise=None, res=None, res_bn=None, want_f32=False, want_pair=True, x_f32_nchw=None, defer_act=False):
        """x: Act (pair required unless strided with x.f32 / x_f32_nchw).  res: FP32 residual (raw) or, with res_bn, the
        pre-BN output of the downsample unit.  Returns (out Act, tape record).  defer_act=True stops after conv + BN stats
        (used by the stem and the head, which fuse their own epilogues)."""
        dev = self.w.device
        if self.stride != 1 or x_f32_nchw is not None:
            src = x_f32_nchw if x_f32_nchw is not None else x.f32
            xin = im2col(src, x_f32_nchw is not None, x.B, x.H, x.W, x.C, self.k, self.stride, self.pad)
            wp, wd = prep_weight_pair(self.w, im2col=True, nsplit=self.ns) if record else (prep_weight(self.w, im2col=True, nsplit=self.ns), None)  # [co, (r,s,c)] = the im2col K order
            kk = 1
        else:
            xin, kk = x, self.k
            if self.glob:
                wp = wd = None  # made from the W_a half of the weight in _glob_forward
            else:
                wp, wd = prep_weight_pair(self.w, nsplit=self.ns) if record else (prep_weight(self.w, nsplit=self.ns), None)
        B, H, W = xin.B, xin.H, xin.W
        P, C = B * H * W, self.cout
        if self.bn is None and self.act == ACT_RELU and noise is None and res is None and not defer_act:
            # Conv1d/Linear + bias + ReLU (per-point MLPs): bias, ReLU and the operand split of the NEXT layer are the
            # GEMM's epilogue; nothing else touches the activation
            out = Act(B, H, W, C)
            if want_f32:
                out.f32 = torch.empty(B, H, W, C, dtype=torch.float32, device=dev)
            if want_pair or record:
                out.pl = empty_planes(B, H, W, C, dev, nsplit=self.ns_out)
            rec = {"bn": None}
            if self.glob:
                wp, wd = self._glob_forward(xin, P, C, record, rec)
                conv_gemm(xin, wp, C, 1, 1, bias=rec["glob"][5], bias_group=self.glob, relu=True, out_f32=out.f32, out_pl=out.pl)
            else:
                conv_gemm(xin, wp, C, kk, kk, bias=self.b, relu=True, out_f32=out.f32, out_pl=out.pl)
            if record:
                rec.update({"xin": xin, "y": None, "noise": None, "kk": kk, "P": P, "HW": H * W, "in_shape": (x.B, x.H, x.W, x.C),
                            "z_hi": out.hi, "wd": wd})
            return out, rec
        y = torch.empty(B, H, W, C, dtype=torch.float32, device=dev)
        st, fin = bn_begin(self.bn, C, training, P, dev) if self.bn is not None else (None, None)
        if fin is not None and FUSE_BN_STATS:  # statistics in the GEMM epilogue, finished by its last CTA
            conv_gemm(xin, wp, C, kk, kk, bias=self.b, out_f32=y, stat_part=stat_scratch(C, dev), fin=fin)
        else:
            conv_gemm(xin, wp, C, kk, kk, bias=self.b, out_f32=y)
            if fin is not None:
                bn_stats(y, P, C, fin)
        rec = {"bn": st}
        if record:
            rec.update({"xin": xin, "y": y, "noise": noise, "kk": kk, "P": P, "HW": H * W, "in_shape": (x.B, x.H, x.W, x.C), "wd": wd,
                        "has_res": res is not None})
        if defer_act:
            return Act(B, H, W, C, y), rec
        if self.bn is None and self.act == ACT_NONE and noise is None and res is None:  # plain linear layer
            out = Act(B, H, W, C, y)
            if want_pair:
                out.pl = empty_planes(B, H, W, C, dev, nsplit=self.ns_out)
                split(y, P, C, out.pl)
            if record:
                rec["y"] = None
            return out, rec
        out = Act(B, H, W, C)
        if want_f32:
            out.f32 = torch.empty(B, H, W, C, dtype=torch.float32, device=dev)
        if want_pair:
            out.pl = empty_planes(B, H, W, C, dev, nsplit=self.ns_out)
        elif record and self.act == ACT_RELU:
            out.pl = empty_planes(B, H, W, C, dev, nsplit=1)  # plane 0 only: the ReLU mask of the backward pass
        bn_act_split(y, P, C, H * W, bn=st, res=res, res_bn=res_bn, act=self.act, prelu=self.prelu, noise=noise, out_f32=out.f32, out_pl=out.pl)
        if record:
            rec["z_hi"] = out.hi
            if self.bn is None and self.act != ACT_PRELU:
                rec["y"] = None  # not needed by the backward of a BN-free ReLU/identity unit
        return out, rec

    def _glob_forward(self, xin, P, C, record, rec):
        """Global-feature unit: per-instance mean of the input rows, bias table W_b mean + b, operand planes of W_a."""
        n, ca, dev = self.glob, xin.C, self.w.device
        if not (self.bn is None and self.act == ACT_RELU and self.k == 1 and self.stride == 1 and xin.f32 is not None and P % n == 0
                and self.w.numel() == C * 2 * ca):
            raise RuntimeError("istnet_b200: global-feature unit needs a Conv1d(2*c -> cout)+ReLU on [rows, c] FP32 input rows")
        nb = P // n
        w2 = self.w.reshape(C, 2 * ca)
        wa, wb = w2[:, :ca].contiguous(), w2[:, ca:].contiguous()
        m = torch.empty(nb, ca, dtype=torch.float32, device=dev)
        _C.call("rows_mean", c_int(nb), c_int(n), c_int(ca), ptr(xin.f32), ptr(m))
        table = torch.empty(nb, C, dtype=torch.float32, device=dev)
        _C.call("heads_linear", c_int(1), c_int(nb), c_int(ca), _ptrs([m]), _ptrs([wb]), _ptrs([self.b]), _ptrs([table]), (ctypes.c_int * 1)(C), c_int(0))
        wp, wd = prep_weight_pair(wa, nsplit=self.ns) if record else (prep_weight(wa, nsplit=self.ns), None)
        rec["glob"] = (m, wb, n, nb, ca, table)
        return wp, wd

    def _glob_data_grads(self, rec, dy, need_dx, grads):
        """Backward of the global-feature unit: S[b] = sum of dy over the rows of instance b is the gradient of the bias table, so
        dW_b = S^T mean, d mean = S W_b, and d f = dy W_a + (d mean)[b] / n — the last term again as a per-instance bias, of the
        data-gradient GEMM."""
        xin, C = rec["xin"], self.cout
        m, wb, n, nb, ca, _ = rec["glob"]
        dev = dy.device
        main = torch.cuda.current_stream()
        side = _side_stream(dev) if (need_dx and WGRAD_SIDE_STREAM) else None
        if side is not None:
            side.wait_stream(main)
        S = torch.empty(nb, C, dtype=torch.float32, device=dev)
        _C.call("rows_group_sum_planes", ptr(dy), c_ll(dy.stride(0)), c_int(dy.shape[0]), c_ll(rec["P"]), c_int(C), c_int(dy.shape[-1]), c_int(n), ptr(S))
        dm = torch.empty(nb, ca, dtype=torch.float32, device=dev)
        dwb = torch.empty(C, ca, dtype=torch.float32, device=dev)
        _C.call("heads_linear_bwd", c_int(1), c_int(nb), c_int(ca), _ptrs([m]), _ptrs([wb]), NULL, _ptrs([S]), _ptrs([dm]), _ptrs([dwb]), NULL,
                (ctypes.c_int * 1)(C), c_int(0))
        with torch.cuda.stream(side if side is not None else main):
            gwa = conv_wgrad(dy, C, xin, 1, 1)
        if side is not None:
            dy.record_stream(side)
            xin.pl.record_stream(side)
            gwa.record_stream(main)
            main.wait_stream(side)  # the two halves of the weight gradient are concatenated on the main stream
        grads[id(self.w)] = torch.cat([gwa.view(C, ca), dwb], 1).reshape(self.w.shape)
        if not need_dx:
            return None
        wd = rec.get("wd")
        dx = torch.empty(xin.B, xin.H, xin.W, ca, dtype=torch.float32, device=dev)
        conv_gemm(Act(xin.B, xin.H, xin.W, C, None, dy), wd, ca, 1, 1, out_f32=dx, bias=dm.mul_(1.0 / n), bias_group=n)
        return dx

    # ---- backward
    def backward(self, rec, dz, dz2=None, need_dx=True, g_out=False, grads=None):
        """dz (+dz2): FP32 gradient w.r.t. the unit output.  Fills grads[param] and returns (dx FP32 or None, g or None)."""
        dy, g = self.act_backward(rec, dz, dz2, g_out, grads)
        return self.data_grads(rec, dy, need_dx, grads), g

    def act_backward(self, rec, dz, dz2, g_out, grads):
        """Activation / BatchNorm backward: FP32 gradient w.r.t. the unit output -> dy operand planes of the convolution's
        backward GEMMs (+ the gradients of bias / BN / PReLU parameters).  Returns (dy planes, g or None)."""
        dev = dz.device
        xin, P, C = rec["xin"], rec["P"], self.cout
        B, H, W = xin.B, xin.H, xin.W
        dy = empty_planes(B, H, W, C, dev, nsplit=NSPLIT_BWD)
        if self.bn is None and self.act == ACT_NONE and rec["noise"] is None:  # plain linear layer
            d = dz if dz2 is None else dz + dz2
            split(d.contiguous(), P, C, dy)
            if self.b is not None:
                grads[id(self.b)] = d.reshape(P, C).sum(0)
            return dy, (d if g_out else None)
        g = torch.empty(B, H, W, C, dtype=torch.float32, device=dev) if g_out else None
        sums = bn_act_bwd(dz, dz2, rec["y"], P, C, rec["HW"], rec["bn"], self.act, self.prelu, rec.get("z_hi"), rec["noise"], dy_pl=dy, g_out=g)
        self.param_grads(rec, sums, grads)
        return dy, g

    def is_bias_relu(self, rec):
        """Conv + bias + ReLU with nothing else (the nn.Conv1d / nn.ReLU stacks): its activation backward can ride in the
        epilogue of the data-gradient GEMM of the layer above (conv_gemm mask_hi)."""
        return self.bn is None and self.act == ACT_RELU and rec.get("noise") is None and rec.get("z_hi") is not None and rec.get("y") is None

    def param_grads(self, rec, sums, grads):
        ws, sg, sgx = sums
        C = self.cout
        if self.bn is not None:
            grads[id(self.bn.weight)] = sgx
            grads[id(self.bn.bias)] = sg
            if self.b is not None:  # a bias feeding a train-mode BatchNorm has an identically zero gradient
                if rec["bn"].batch:
                    grads[id(self.b)] = torch.zeros_like(self.b)
                else:  # running statistics: d bias = sum_p dy = gamma*invstd*sum g
                    grads[id(self.b)] = (rec["bn"].gamma * rec["bn"].invstd * sg).detach()
        elif self.b is not None:
            grads[id(self.b)] = sg
        if self.act == ACT_PRELU:
            grads[id(self.prelu)] = ws[2 * C : 3 * C].sum().float().reshape(1)

    def is_bn_relu(self, rec):
        """Conv + train-mode BatchNorm + ReLU, no residual / Dropout2d: its BN-backward reduction can ride in the epilogue of the
        data-gradient GEMM of the layer above (conv_gemm mask_hi + stat_y)."""
        st = rec.get("bn")
        return (self.bn is not None and st is not None and st.batch and self.act == ACT_RELU and rec.get("noise") is None and not rec.get("has_res")
                and rec.get("z_hi") is not None and rec.get("y") is not None)

    def data_grads(self, rec, dy, need_dx, grads, below=None, below_f32=False):
        """Weight gradient (side stream) and data gradient of the convolution.  below = (unit, rec) of the bias+ReLU layer that
        produced this unit's input: the data-gradient GEMM then applies that layer's ReLU mask in its epilogue and returns
        ITS dy operand planes (and fills its bias gradient) instead of the FP32 dx."""
        if rec.get("glob") is not None:
            assert below is None
            return self._glob_data_grads(rec, dy, need_dx, grads)
        xin, kk, C = rec["xin"], rec["kk"], self.cout
        # weight gradient and data gradient are independent: wgrad goes to a side stream and overlaps the dgrad GEMM
        main = torch.cuda.current_stream()
        side = _side_stream(dy.device) if (need_dx and WGRAD_SIDE_STREAM) else None
        if side is not None:
            side.wait_stream(main)
        with torch.cuda.stream(side if side is not None else main):
            gw = conv_wgrad(dy, C, xin, kk, kk)
            if kk != self.k or self.stride != 1:  # im2col path: [co, (r,s,c)] -> [co, c, r, s]
                cin = self.w.shape[1]
                gw = gw.reshape(C, self.k, self.k, cin).permute(0, 3, 1, 2).contiguous()
        grads[id(self.w)] = gw.reshape(self.w.shape)
        if side is not None:
            dy.record_stream(side)
            xin.pl.record_stream(side)
            gw.record_stream(main)
            _PENDING_JOINS.append((main, side))
        if not need_dx:
            return None
        dyA = Act(xin.B, xin.H, xin.W, C, None, dy)
        if kk != self.k or self.stride != 1:
            wd = rec.get("wd")
            if wd is None or wd.shape[0] != dy.shape[0]:
                wd = prep_weight(self.w, transpose=True, nsplit=dy.shape[0], im2col=True)
            dcol = torch.empty(xin.B, xin.H, xin.W, xin.C, dtype=torch.float32, device=dy.device)
            conv_gemm(dyA, wd, xin.C, 1, 1, out_f32=dcol)
            b, h, w, c = rec["in_shape"]
            return col2im(dcol, b, h, w, c, self.k, self.stride, self.pad)
        wd = rec.get("wd")
        if wd is None or wd.shape[0] != dy.shape[0]:
            wd = prep_weight(self.w, transpose=True, nsplit=dy.shape[0])
        if below is not None and below[0].bn is not None:
            # conv+BN+ReLU below: g = dx*mask (FP32) and the partials of sum g, sum g*y come out of this GEMM; finalize; apply pass
            bu, brec = below
            st, Cb, dev = brec["bn"], xin.C, dy.device
            g = torch.empty(xin.B, xin.H, xin.W, Cb, dtype=torch.float32, device=dev)
            part = stat_scratch(Cb, dev)
            G = conv_gemm(dyA, wd, Cb, kk, kk, out_f32=g, stat_part=part, mask_hi=brec["z_hi"], stat_y=brec["y"])
            ws = torch.empty(3 * Cb, dtype=torch.float64, device=dev)
            sg = torch.empty(Cb, dtype=torch.float32, device=dev)
            sgx = torch.empty(Cb, dtype=torch.float32, device=dev)
            _C.call("bn_bwd_finalize_gy", ptr(part), c_int(G), c_int(Cb), ptr(st.mean), ptr(st.invstd), ptr(ws), ptr(sg), ptr(sgx))
            grads[id(bu.bn.weight)], grads[id(bu.bn.bias)] = sgx, sg
            if bu.b is not None:
                grads[id(bu.b)] = torch.zeros_like(bu.b)
            dyb = None if below_f32 else empty_planes(xin.B, xin.H, xin.W, Cb, dev, nsplit=NSPLIT_BWD)
            dyf = torch.empty_like(g) if below_f32 else None
            zh = brec["z_hi"]
            _C.call("bn_bwd_apply", ptr(g), ptr(brec["y"]), c_ll(xin.P), c_int(Cb), ptr(st.mean), ptr(st.invstd), ptr(st.gamma), ptr(st.beta), c_int(1),
                    ptr(zh), c_int(zh.shape[-1]), ptr(ws), *_pl_args(dyb), c_int(dyb.shape[-1] if dyb is not None else 0), _p(dyf))
            return dyf if below_f32 else dyb
        if below is not None:
            bu, brec = below
            dyb = empty_planes(xin.B, xin.H, xin.W, xin.C, dy.device, nsplit=NSPLIT_BWD)
            if bu.b is not None:  # bias gradient = column sums of the masked dx: statistics epilogue, finished by the GEMM's last CTA
                gb = torch.empty(xin.C, dtype=torch.float32, device=dy.device)
                fin = _C.Fin()
                fin.kind, fin.tickets, fin.sum_f32 = _C.FIN_COLSUM, _C.tickets(dy.device).value, gb.data_ptr()
                conv_gemm(dyA, wd, xin.C, kk, kk, out_pl=dyb, stat_part=stat_scratch(xin.C, dy.device), mask_hi=brec["z_hi"], fin=fin)
                grads[id(bu.b)] = gb
            else:
                conv_gemm(dyA, wd, xin.C, kk, kk, out_pl=dyb, mask_hi=brec["z_hi"])
            return dyb
        dx = torch.empty(xin.B, xin.H, xin.W, xin.C, dtype=torch.float32, device=dy.device)
        conv_gemm(dyA, wd, xin.C, kk, kk, out_f32=dx)
        return dx
