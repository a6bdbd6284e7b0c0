"""Backward pass of the engine-backed UNets: the hand-written adjoint of models/engine.py's forward.

The reference trains through autograd (`loss.backward()` in scripts/train_ddpm.py:171-192 over models/unet.py).
Here `UNetFunction` is ONE autograd node around the whole network: its forward runs the forward kernels in training
mode (dropout on, every composite op appends a record to `engine.tape`), its backward replays the tape in reverse
and issues only libb200diff kernels:
  * conv data gradients  = the forward implicit-GEMM kernel (K1) on spatially flipped, channel-transposed weights;
  * conv weight gradients = b200_conv2d_wgrad (tcgen05, contraction over pixels via MN-major TMA tiles), written
    straight into the reference-layout (OIHW) gradient tensors;
  * GroupNorm / AdaGN / SiLU / dropout / resample adjoints = b200_groupnorm_bwd (two streaming passes);
  * attention adjoint = batched tcgen05 GEMMs (b200_gemm_batched) around the row-softmax kernels;
  * embedding path adjoint = b200_time_embed_bwd + GEMMs.
Gradients w.r.t. parameters are returned to autograd (so `.grad` accumulation, DDP hooks, torch optimizers and
`clip_grad_norm_` behave exactly as with the reference); they are views into one flat fp32 buffer per backward.
"""
from typing import Dict

import torch
import torch.nn as nn

import b200diff as K
from models.engine import Act, Engine


class UNetFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, model, args, kwargs, *params):
        eng = model.engine
        eng.tape = []
        eng.new_dropout_base()
        try:
            out = model._forward_impl(*args, **kwargs)
            ctx.tape = eng.tape
            ctx.arena_key = eng._cur_key               # (input shape, lane) whose arena buffers the tape refers to
            ctx.arena_gen = eng._fwd_gen.get(ctx.arena_key)
            eng.last_tape = eng.tape     # introspection (tests read the dropout seeds from it)
        finally:
            eng.tape = None
            eng.pingpong = False     # inference-only buffer sharing (Engine.buf) must not leak into the backward's requests
        ctx.model = model
        ctx.params = params
        return out

    @staticmethod
    def backward(ctx, dout):
        eng = ctx.model.engine
        if ctx.tape is None:
            raise RuntimeError('b200diff UNet: backward called twice on the same forward (the tape is consumed; '
                               'retain_graph is not supported)')
        if ctx.arena_gen is not None and eng._fwd_gen.get(ctx.arena_key) != ctx.arena_gen:
            raise RuntimeError(
                'b200diff UNet: another forward with the same input shape ran between this training forward and its '
                'backward.  The tape refers to the engine\'s shared activation arena (no saved copies), so the saved '
                'activations were overwritten; run backward() before the next forward of this model (evaluation / '
                'sampling forwards included), or use a second model instance.')
        grads = run_backward(ctx.model.engine, ctx.tape, dout.contiguous(), ctx.params)
        ctx.tape = None
        return (None, None, None) + tuple(grads)


class MSELossFunction(torch.autograd.Function):
    """F.mse_loss(pred, target) (mean reduction, diffusions/ddpm.py:136-138) as one kernel each way."""

    @staticmethod
    def forward(ctx, pred, target):
        pred, target = pred.contiguous(), target.contiguous().to(pred.dtype)
        loss = torch.empty(1, dtype=torch.float32, device=pred.device)
        K.mse_loss(pred, target, loss)
        ctx.save_for_backward(pred, target)
        return loss.view(())

    @staticmethod
    def backward(ctx, g):
        pred, target = ctx.saved_tensors
        da = torch.empty_like(pred)
        K.mse_loss_grad(pred, target, g.reshape(1).float().contiguous(), da)
        return da, None


def mse_loss(pred, target):
    if pred.shape != target.shape:
        raise RuntimeError(f'mse_loss: shape mismatch {tuple(pred.shape)} vs {tuple(target.shape)}')
    return MSELossFunction.apply(pred, target)


def _record_params(e):
    """Parameters whose gradients a tape record's backward writes: those of the modules the record references.  The
    per-block embedding projection (`emb_linear`) belongs to the embedding record, which runs last."""
    out = []
    for k, v in e.items():
        if k == 'emb_linear':
            continue
        if k == 'linears' and e.get('kind') == 'embed':
            # all projection weights side by side, then all biases: _embed_bwd writes each group with ONE launch
            out += [l.weight for l in v] + [l.bias for l in v]
            continue
        if k == 'mods' and e.get('kind') == 'attn' and isinstance(v, tuple):
            # separate q, k, v convs: their weights (and biases) side by side, so that _attn_bwd computes the three
            # weight gradients as ONE GEMM over the [dq | dk | dv] columns and the bias gradients as one column sum
            q, kk, vv, proj = v
            out += [q.weight, kk.weight, vv.weight, q.bias, kk.bias, vv.bias] + list(proj.parameters())
            continue
        vs = v if isinstance(v, (tuple, list)) else (v,)
        for m in vs:
            if isinstance(m, nn.Module):
                out += list(m.parameters())
    return out


def _completion_layout(tape, params):
    """Order in which parameter gradients become FINAL during `run_backward` (head first, the embedding path last) and,
    per tape position, how many parameters are complete once that record's backward has been issued.  The flat gradient
    buffer is laid out in this order so that finished prefixes of it can be all-reduced while the rest of the backward
    still runs (what DistributedDataParallel's buckets do for the reference, scripts/train_ddpm.py:166,180-184)."""
    known = {id(p) for p in params}
    order, seen, done_after = [], set(), {}
    recs = [e for e in reversed(tape) if e['kind'] != 'embed'] + [e for e in tape if e['kind'] == 'embed']
    for i, e in enumerate(recs):
        for p in _record_params(e):
            if id(p) in known and id(p) not in seen:
                seen.add(id(p))
                order.append(p)
        done_after[id(e)] = len(order)
    order += [p for p in params if id(p) not in seen]      # parameters this forward did not use: their gradient stays 0
    return order, done_after


class _Grads:
    """Per-parameter fp32 gradient views into one flat, zero-initialised buffer (the kernels accumulate), laid out in
    gradient-completion order (see _completion_layout)."""

    def __init__(self, eng: Engine, params, tape=None):
        total = sum(p.numel() for p in params)
        # One persistent buffer per engine, zeroed per backward: stable gradient addresses across steps (optimizer
        # chunk tables, CUDA-graph capture, NCCL buffer registration).  autograd keeps the returned views as the
        # parameters' .grad, so when a .grad still aliases the buffer (gradient accumulation over micro-batches: a second
        # backward before zero_grad) this backward gets a fresh buffer instead and autograd adds it to the first.
        flat = getattr(eng, '_grad_flat', None)
        aliased = False
        if flat is not None and flat.numel() == total and flat.device == eng.device:
            lo, hi = flat.data_ptr(), flat.data_ptr() + 4 * total
            aliased = any(p.grad is not None and lo <= p.grad.data_ptr() < hi for p in params)
        if flat is None or flat.numel() != total or flat.device != eng.device:
            flat = eng._grad_flat = torch.zeros(total, dtype=torch.float32, device=eng.device)
        elif aliased:
            flat = torch.zeros(total, dtype=torch.float32, device=eng.device)
        else:
            flat.zero_()
        self.flat = flat
        self.aliased = aliased
        # the layout depends on the network structure only: computed from the first tape, reused afterwards (a forward
        # that takes another path -- e.g. with / without class labels -- keeps the addresses; its records simply
        # complete in a slightly different order, which `ready_prefix` accounts for per backward)
        layout = getattr(eng, '_grad_layout', None)
        if layout is None or layout[0] != tuple(id(p) for p in params):
            order = _completion_layout(tape, params)[0] if tape is not None else list(params)
            layout = eng._grad_layout = (tuple(id(p) for p in params), order)
        self.order = layout[1]
        self.views: Dict[int, torch.Tensor] = {}
        self.offset: Dict[int, int] = {}
        off = 0
        for p in self.order:
            self.views[id(p)] = self.flat[off:off + p.numel()].view(p.shape)
            self.offset[id(p)] = off
            off += p.numel()
        self.params = params

    def __call__(self, p):
        return self.views[id(p)]

    def as_tuple(self):
        return [self.views[id(p)] if p.requires_grad else None for p in self.params]


class _OverlappedAllReduce:
    """Data-parallel gradient averaging overlapped with the backward: as soon as the records issued so far have completed
    a prefix of the flat gradient buffer that is `bucket_bytes` longer than what was already sent, that slice is
    all-reduced asynchronously (NCCL's stream waits for the producing kernels, the compute stream carries on with the
    remaining weight gradients); `finish()` sends the tail and makes the compute stream wait for all of them."""

    def __init__(self, G: _Grads, tape, bucket_bytes: int):
        import torch.distributed as dist
        self.dist, self.G = dist, G
        self.world = dist.get_world_size()
        self.bucket = max(1, bucket_bytes // 4)
        self.sent = 0
        self.works = []
        self.n_buckets = 0
        self.avg = dist.get_backend() == 'nccl'
        # elements of the flat buffer that are final after each record: the longest prefix of the layout whose parameters
        # all belong to records already issued in THIS backward
        self.pending = set()
        for e in tape:
            for p in _record_params(e):
                self.pending.add(id(p))
        self.prefix_param = 0

    def _ready(self):
        order = self.G.order
        i = self.prefix_param
        while i < len(order) and id(order[i]) not in self.pending:
            i += 1
        self.prefix_param = i
        return self.G.offset[id(order[i])] if i < len(order) else self.G.flat.numel()

    def after_record(self, e):
        for p in _record_params(e):
            self.pending.discard(id(p))
        ready = self._ready()
        if ready - self.sent >= self.bucket:
            self._send(ready)

    def _send(self, hi):
        if hi <= self.sent:
            return
        sl = self.G.flat[self.sent:hi]
        op = self.dist.ReduceOp.AVG if self.avg else self.dist.ReduceOp.SUM
        self.works.append((self.dist.all_reduce(sl, op=op, async_op=True), sl))
        self.sent = hi
        self.n_buckets += 1

    def finish(self):
        self.pending.clear()
        self._send(self.G.flat.numel())
        for w, sl in self.works:
            w.wait()
            if not self.avg:
                sl.div_(self.world)


# ------------------------------------------------------------------------------------------------------
# helpers
# ------------------------------------------------------------------------------------------------------
def _grad_slot(eng: Engine, act: Act):
    """(fp32 gradient buffer of `act`, accumulate?) -- the first contribution stores, later ones accumulate."""
    if act.g is None:
        act.g = eng.buf(('grad', act.t.data_ptr()), tuple(act.t.shape), torch.float32)
        return act.g, False
    return act.g, True


def _w_dgrad(eng: Engine, tag, conv: nn.Conv2d):
    """Packed weights of the data-gradient conv: spatially flipped taps, in/out channels swapped."""
    return eng.w_dgrad(tag, conv)


def _cast_out_grad(eng: Engine, tag, out: Act, bias_grad):
    """fp32 gradient of a conv output -> bf16 tensor-core operand, bias gradient = its column sums."""
    B, H, W, C = out.B, out.H, out.W, out.C
    dob = eng.buf(tag + '.dOb', (B, H, W, C), torch.bfloat16)
    K.cast_bf16_colsum(out.g, dob, bias_grad, B * H * W, C)
    return dob


def _wgrad(eng: Engine, *args, **kw):
    """b200_conv2d_wgrad with the engine's shared split-K workspace (two-phase, atomics-free reduction)."""
    ws = eng._arena.get('wgrad_scratch')
    if ws is None:
        ws = eng._arena['wgrad_scratch'] = torch.empty(64 << 20, dtype=torch.float32, device=eng.device)   # 256 MB
    return K.conv2d_wgrad(*args, scratch=ws, **kw)


def _sums(eng: Engine, B, C):
    return eng.buf('gn_bwd_sums', (B, 8, C), torch.float32)


# ------------------------------------------------------------------------------------------------------
# composite adjoints
# ------------------------------------------------------------------------------------------------------
def _res_bwd(eng: Engine, e, G: _Grads):
    tag, x, skip, out = e['tag'], e['x'], e['skip'], e['out']
    conv1, conv2, norm1, norm2, sc = e['conv1'], e['conv2'], e['norm1'], e['norm2'], e['shortcut']
    B, H, W = x.B, x.H, x.W
    Ho, Wo, Cout = out.H, out.W, out.C
    Cin = x.C + (skip.C if skip is not None else 0)
    resample = e['resample']
    t3, t1 = K.taps_3x3_s1(), K.taps_1x1()
    if out.g is None:
        raise RuntimeError(f'{tag}: no gradient reached this block')

    # conv2 (+ shortcut): bias, weights, data
    dob = _cast_out_grad(eng, tag + '.c2', out, G(conv2.bias))
    _wgrad(eng, dob, Cout, e['a2'], (Cout, Ho, Wo, 1), B, Ho, Wo, Cout, Cout, t3[0], G(conv2.weight))
    addend = None
    if sc is not None:
        G(sc.bias).copy_(G(conv2.bias))
        k = sc.kernel_size[0]
        _wgrad(eng, dob, Cout, e['raw'], (Cin, Ho, Wo, 1), B, Ho, Wo, Cout, Cin, (t1 if k == 1 else t3)[0],
                       G(sc.weight))
        addend = eng.buf(tag + '.dcat', (B, H, W, Cin), torch.float32)
        K.conv2d(dob, _w_dgrad(eng, tag + '.sc', sc), Cin, B, H, W, t1 if k == 1 else t3, a0_geom=(Cout, H, W, 1),
                 out=addend)
    da2 = eng.buf(tag + '.dA2', (B, Ho, Wo, Cout), torch.bfloat16)
    K.conv2d(dob, _w_dgrad(eng, tag + '.c2', conv2), Cout, B, Ho, Wo, t3, a0_geom=(Cout, Ho, Wo, 1), out=da2,
             out_mode=K.OUT_BF16_NHWC)

    # GroupNorm 2 (+ AdaGN, SiLU, dropout) -> dH; conv1 bias and embedding-projection gradients ride along
    h = e['h']
    dh = eng.buf(tag + '.dH', (B, Ho, Wo, Cout), torch.bfloat16)
    emb, off, total = e['emb'], e['emb_off'], eng.d_emb.shape[1]
    kw = {}
    if e['scale_shift']:
        kw = dict(scale=emb[:, off:], shift=emb[:, off + Cout:], ss_ld=e['emb_ld'], dscale=eng.d_emb[:, off:],
                  dshift=eng.d_emb[:, off + Cout:], dss_ld=total)
    else:
        kw = dict(dx_rowsum=eng.d_emb[:, off:], dx_rowsum_ld=total)
    K.groupnorm_bwd(da2, h.t, Cout, h.stats, None, 0, None, B, Ho * Wo, Wo, norm2.num_groups, norm2.weight, norm2.bias,
                    norm2.eps, _sums(eng, B, Cout), silu=True, drop_p=e['drop_p'], drop_seed=e['drop_seed'],
                    drop_seed_dev=eng.seed_dev if e['drop_p'] > 0 else None, dx_bf16=dh,
                    dx_colsum=G(conv1.bias), dgamma=G(norm2.weight), dbeta=G(norm2.bias), **kw)

    # conv1: weights, data
    _wgrad(eng, dh, Cout, e['a1'], (Cin, Ho, Wo, 1), B, Ho, Wo, Cout, Cin, t3[0], G(conv1.weight))
    da1 = eng.buf(tag + '.dA1', (B, Ho, Wo, Cin), torch.bfloat16)
    K.conv2d(dh, _w_dgrad(eng, tag + '.c1', conv1), Cin, B, Ho, Wo, t3, a0_geom=(Cout, Ho, Wo, 1), out=da1,
             out_mode=K.OUT_BF16_NHWC)

    # identity shortcut of a resampling block: the residual gradient reaches x through the resample adjoint
    if sc is None and resample:
        gx, acc = _grad_slot(eng, x)
        if resample == 1:    # forward avg-pooled x: spread a quarter of the gradient over each 2x2 block
            K.resample_f32(out.g, gx, B, Ho, Wo, Cout, 2, scale=0.25, accumulate=acc)
        else:                # forward replicated x 2x2: sum the block
            K.resample_f32(out.g, gx, B, Ho, Wo, Cout, 1, scale=4.0, accumulate=acc)
    elif sc is None:
        addend = out.g

    # GroupNorm 1 (+ SiLU, resample) over cat(x, skip) -> gradients of x and of the skip connection
    gx, accx = _grad_slot(eng, x)
    gs, accs = _grad_slot(eng, skip) if skip is not None else (None, False)
    K.groupnorm_bwd(da1, x.t, x.C, x.stats, None if skip is None else skip.t, 0 if skip is None else skip.C,
                    None if skip is None else skip.stats, B, H * W, W, norm1.num_groups, norm1.weight, norm1.bias,
                    norm1.eps, _sums(eng, B, Cin), silu=True, resample=resample, dx0=gx, dx0_acc=accx, dx1=gs,
                    dx1_acc=accs, addend=addend, dgamma=G(norm1.weight), dbeta=G(norm1.bias))


def _attn_bwd(eng: Engine, e, G: _Grads):
    tag, x, out, norm = e['tag'], e['x'], e['out'], e['norm']
    B, H, W, C = x.B, x.H, x.W, x.C
    targets, proj = _attn_targets(eng, e, G, C)
    T, heads, scale = H * W, e['heads'], e['scale']
    d = C // heads
    t1 = K.taps_1x1()
    qk, vt, o = e['qk'], e['vt'], e['o']
    bf = torch.bfloat16

    dob = _cast_out_grad(eng, tag + '.proj', out, G(proj.bias))
    _wgrad(eng, dob, C, o, (C, H, W, 1), B, H, W, C, C, t1[0], G(proj.weight))
    do = eng.buf(tag + '.dO', (B, T, C), bf)
    K.conv2d(dob, _w_dgrad(eng, tag + '.proj', proj), C, B, H, W, t1, a0_geom=(C, H, W, 1), out=do,
             out_mode=K.OUT_BF16_NHWC)

    # attention core: recompute P, then dV = P^T dO, dP = dO V^T, dS = softmax', dQ = dS K, dK = dS^T Q
    dqkv = eng.buf(tag + '.dqkv', (B, T, 3 * C), bf)          # [dq | dk | dv] side by side
    dqk, dv = dqkv[:, :, :2 * C], dqkv[:, :, 2 * C:]
    if e.get('lse') is not None:
        K.attention_bwd(qk, vt, o, do, e['lse'], dqk, dv, B, T, heads, d, scale)      # one launch, scores stay on chip
    else:
        _attn_core_bwd_gemms(eng, qk, vt, do, dqk, dv, B, T, C, heads, d, scale)

    # q, k, v 1x1 convs: biases, weights, data (one GEMM over the concatenated [dq | dk | dv] channels).
    # `targets` maps row ranges of dq / dk / dv to rows of the parameters' gradients: one range each for separate
    # q, k, v convs; one per head for ADM's fused, head-interleaved qkv Conv1d.
    n = e['n']
    mods = e['mods']
    fused_qkv = False
    if isinstance(mods, tuple):
        q, k, v = mods[:3]
        w0, b0 = G.offset[id(q.weight)], G.offset[id(q.bias)]
        fused_qkv = (G.offset[id(k.weight)] == w0 + C * C and G.offset[id(v.weight)] == w0 + 2 * C * C and
                     G.offset[id(k.bias)] == b0 + C and G.offset[id(v.bias)] == b0 + 2 * C)
    if fused_qkv:    # gradient layout keeps q, k, v weights / biases adjacent (_record_params): one GEMM, one column sum
        K.colsum_bf16(dqkv, G.flat[b0:b0 + 3 * C], B * T, 3 * C, 0, 3 * C)
        _wgrad(eng, dqkv, 3 * C, n, (C, H, W, 1), B, H, W, 3 * C, C, t1[0], G.flat[w0:w0 + 3 * C * C].view(3 * C, C, 1, 1))
    else:
        for which, c_off in (('q', 0), ('k', C), ('v', 2 * C)):
            for (r0, nrows, wg, bg) in targets[which]:
                K.colsum_bf16(dqkv, bg, B * T, 3 * C, c_off + r0, nrows)
                _wgrad(eng, dqkv, 3 * C, n, (C, H, W, 1), B, H, W, nrows, C, t1[0], wg, dy_c0=c_off + r0)
    wd = targets['wd']()
    dn = eng.buf(tag + '.dN', (B, H, W, C), bf)
    K.conv2d(dqkv, wd, C, B, H, W, t1, a0_geom=(3 * C, H, W, 1), out=dn, out_mode=K.OUT_BF16_NHWC)

    gx, acc = _grad_slot(eng, x)
    K.groupnorm_bwd(dn, x.t, C, x.stats, None, 0, None, B, T, W, norm.num_groups, norm.weight, norm.bias, norm.eps,
                    _sums(eng, B, C), silu=False, dx0=gx, dx0_acc=acc, addend=out.g, dgamma=G(norm.weight),
                    dbeta=G(norm.bias))


def _attn_core_bwd_gemms(eng: Engine, qk, vt, do, dqk, dv, B, T, C, heads, d, scale):
    """Attention-core adjoint for shapes b200_attention_bwd does not cover (head dim != 64 or T > 256): five batched
    tensor-core GEMMs around the row-softmax kernels, [B*h, T, T] workspaces in HBM."""
    bf = torch.bfloat16
    G_ = B * heads
    S = eng.buf('attn_ws.S', (G_, T, T), torch.float32)
    P = eng.buf('attn_ws.P', (G_, T, T), bf)
    dP = eng.buf('attn_ws.dP', (G_, T, T), torch.float32)
    dS = eng.buf('attn_ws.dS', (G_, T, T), bf)
    qop = (qk, T, 2 * C, dict(col_base=0, col_head=d))
    kop = (qk, T, 2 * C, dict(col_base=C, col_head=d))
    grid = dict(batch=B, heads=heads)
    sq = dict(out_ld=T, out_batch_stride=heads * T * T, out_head_stride=T * T)
    K.gemm_batched(qop, kop, S, T, T, d, **grid, **sq)
    K.softmax_rows(S, P, G_ * T, T, scale)
    vop = (vt, d, T, dict(per_head_batch=True, mn_major=True))                 # [B*h][d][T]: rows = channel (K), cols = key
    K.gemm_batched((do, T, C, dict(col_head=d)), vop, dP, T, T, d, **grid, **sq)
    K.softmax_bwd_rows(P, dP, dS, G_ * T, T, scale)
    dsop = (dS, T, T, dict(per_head_batch=True))
    K.gemm_batched(dsop, (qk, T, 2 * C, dict(col_base=C, col_head=d, mn_major=True)), dqk, T, d, T, **grid,
                   out_ld=dqk.stride(1), out_batch_stride=dqk.stride(0), out_head_stride=d)
    dsop_t = (dS, T, T, dict(per_head_batch=True, mn_major=True))
    K.gemm_batched(dsop_t, (qk, T, 2 * C, dict(col_base=0, col_head=d, mn_major=True)), dqk[:, :, C:], T, d, T, **grid,
                   out_ld=dqk.stride(1), out_batch_stride=dqk.stride(0), out_head_stride=d)
    K.gemm_batched((P, T, T, dict(per_head_batch=True, mn_major=True)), (do, T, C, dict(col_head=d, mn_major=True)), dv,
                   T, d, T, **grid, out_ld=dv.stride(1), out_batch_stride=dv.stride(0), out_head_stride=d)


def _attn_targets(eng: Engine, e, G: _Grads, C: int):
    """({'q'|'k'|'v': [(first row, rows, weight-grad view [rows, C], bias-grad view [rows])], 'wd': maker of the
    [C][3C] = [Wq^T | Wk^T | Wv^T] data-gradient operand}, output projection module) of an attention block."""
    tag, mods = e['tag'], e['mods']
    bf = torch.bfloat16
    if isinstance(mods, tuple):                       # separate q, k, v, proj convs (models/modules.py, pesser)
        q, k, v, proj = mods

        def wd():
            hit = eng._pt.get(('dgrad', tag + '.qkv'))
            if hit is not None:
                return hit[0]
            w = torch.empty((C, 3 * C), dtype=bf, device=eng.device)
            return eng.pack_table(('dgrad', tag + '.qkv'), w, [K.pack_entry_bytes(m.weight, w, C, C, 1, 1, col0=i * C, ld=3 * C)
                                                               for i, m in enumerate((q, k, v))])
        tg = {n: [(0, C, G(m.weight).view(C, C), G(m.bias))] for n, m in (('q', q), ('k', k), ('v', v))}
        tg['wd'] = wd
        return tg, proj
    blk = mods                                        # ADM AttentionBlock: fused qkv Conv1d [3C, C, 1]
    Hh = blk.num_heads
    d = C // Hh
    wg, bg = G(blk.qkv.weight).view(3 * C, C), G(blk.qkv.bias)
    tg = {}
    for j, name in enumerate(('q', 'k', 'v')):
        if blk.use_new_attention_order:               # rows (j*H + h)*d + i: one contiguous range per projection
            tg[name] = [(0, C, wg[j * C:(j + 1) * C], bg[j * C:(j + 1) * C])]
        else:                                         # legacy rows (h*3 + j)*d + i: one range per head
            tg[name] = [(h * d, d, wg[(h * 3 + j) * d:(h * 3 + j + 1) * d], bg[(h * 3 + j) * d:(h * 3 + j + 1) * d])
                        for h in range(Hh)]

    def wd():
        def make():
            wqk, _, wv, _, _, _ = blk.packed_weights()                   # [2C, C] (q heads | k heads), [C, C]
            return torch.cat([wqk[:C].t(), wqk[C:].t(), wv.t()], dim=1).contiguous()
        return eng.packed(('dgrad', tag + '.qkv'), make)
    tg['wd'] = wd
    return tg, blk.proj_out


def _s2_dgrad_plan(conv: nn.Conv2d, pad_lo: int):
    """Transposed 3x3 stride-2 conv as four 2x2-tap phase convolutions over dY (output pixel (2i+a, 2j+b)):
    tap r contributes to parity a iff (a - r + pad_lo) is even, reading dY row i + (a - r + pad_lo) / 2."""
    w = conv.weight.detach().float()            # [Cout, Cin, 3, 3]
    co, ci = w.shape[0], w.shape[1]

    def axis(a):
        items = [(r, (a - r + pad_lo) // 2) for r in range(3) if (a - r + pad_lo) % 2 == 0]
        return items + [(None, 0)] * (2 - len(items))
    taps, mats = [], []
    for a in range(2):
        for b in range(2):
            ph_taps, cols = [], []
            for (r, dh) in axis(a):
                for (s_, dw) in axis(b):
                    ph_taps.append((dw, dh, 0))
                    cols.append(torch.zeros(ci, co, device=w.device) if r is None or s_ is None
                                else w[:, :, r, s_].t())
            taps.append(ph_taps)
            mats.append(torch.cat(cols, dim=1))           # [Cin][4*Cout]
    return taps, torch.cat(mats, dim=0).to(torch.bfloat16).contiguous()


def _down_bwd(eng: Engine, e, G: _Grads):
    tag, x, out, conv = e['tag'], e['x'], e['out'], e['conv']
    B, H, W, C = x.B, x.H, x.W, x.C
    Ho, Wo, Cout = out.H, out.W, out.C
    dob = _cast_out_grad(eng, tag, out, G(conv.bias))
    _wgrad(eng, dob, Cout, e['planes'], (C, Ho, Wo, 4), B, Ho, Wo, Cout, C, K.taps_3x3_s2(e['pad_lo'])[0],
                   G(conv.weight))
    taps, wd = eng.packed(('dgrad_s2', tag), lambda: _s2_dgrad_plan(conv, e['pad_lo']))
    gx, acc = _grad_slot(eng, x)
    K.conv2d(dob, wd, C, B, Ho, Wo, taps, a0_geom=(Cout, Ho, Wo, 1), out=gx, w_rows_per_phase=C,
             residual=gx if acc else None, res_ld=C)


def _up_bwd(eng: Engine, e, G: _Grads):
    tag, x, out, conv = e['tag'], e['x'], e['out'], e['conv']
    B, H, W, C = x.B, x.H, x.W, x.C
    Ho, Wo, Cout = out.H, out.W, out.C
    t3 = K.taps_3x3_s1()
    dob = _cast_out_grad(eng, tag, out, G(conv.bias))
    _wgrad(eng, dob, Cout, e['ub'], (C, Ho, Wo, 1), B, Ho, Wo, Cout, C, t3[0], G(conv.weight))
    du = eng.buf(tag + '.dU', (B, Ho, Wo, C), torch.float32)
    K.conv2d(dob, _w_dgrad(eng, tag, conv), C, B, Ho, Wo, t3, a0_geom=(Cout, Ho, Wo, 1), out=du)
    gx, acc = _grad_slot(eng, x)
    K.resample_f32(du, gx, B, Ho, Wo, C, 1, scale=4.0, accumulate=acc)     # adjoint of nearest 2x = 2x2 sum


def _head_bwd(eng: Engine, e, G: _Grads, dout):
    tag, x, norm, conv = e['tag'], e['x'], e['norm'], e['conv']
    B, H, W, C = x.B, x.H, x.W, x.C
    Co = conv.out_channels
    t3 = K.taps_3x3_s1()
    dob = eng.buf(tag + '.dOb', (B, H, W, 64), torch.bfloat16)
    K.nchw_to_nhwc_pad_bf16(dout, dob, G(conv.bias), B, Co, H * W, 64)
    _wgrad(eng, dob, 64, e['a'], (C, H, W, 1), B, H, W, Co, C, t3[0], G(conv.weight))

    def make():   # data-gradient weights with the output channels zero-padded to the 64-channel operand
        w = conv.weight.detach().flip(2, 3).transpose(0, 1)          # [C, Co, 3, 3]
        wp = torch.zeros(C, 64, 3, 3, device=w.device, dtype=w.dtype)
        wp[:, :Co] = w
        return K.pack_weight(wp)
    da = eng.buf(tag + '.dA', (B, H, W, C), torch.bfloat16)
    K.conv2d(dob, eng.packed(('dgrad', tag), make), C, B, H, W, t3, a0_geom=(64, H, W, 1), out=da,
             out_mode=K.OUT_BF16_NHWC)
    gx, acc = _grad_slot(eng, x)
    K.groupnorm_bwd(da, x.t, C, x.stats, None, 0, None, B, H * W, W, norm.num_groups, norm.weight, norm.bias, norm.eps,
                    _sums(eng, B, C), silu=True, dx0=gx, dx0_acc=acc, dgamma=G(norm.weight), dbeta=G(norm.bias))


def _first_bwd(eng: Engine, e, G: _Grads):
    tag, X, out, conv = e['tag'], e['X'], e['out'], e['conv']
    B, Ci, H, W = X.shape
    Co = conv.out_channels
    if out.g is None:
        return
    dob = _cast_out_grad(eng, tag, out, G(conv.bias))
    xp = eng.buf(tag + '.xpad', (B, H, W, 64), torch.bfloat16)
    K.nchw_to_nhwc_pad_bf16(X, xp, None, B, Ci, H * W, 64)
    _wgrad(eng, dob, Co, xp, (64, H, W, 1), B, H, W, Co, Ci, K.taps_3x3_s1()[0], G(conv.weight))


def _embed_bwd(eng: Engine, e, G: _Grads):
    rows, E, total = e['rows'], e['E'], e['total']
    rp = (rows + 7) // 8 * 8
    bf = torch.bfloat16
    dev = eng.device
    key = ('embed_bwd_ws', rp, total, E)
    ws = eng._arena.get(key)
    if ws is None:   # zero-initialised once: the padding rows (rows..rp) must stay zero
        dim = e['pos_emb'].dim
        ws = dict(dproj=torch.zeros(rp, total, dtype=bf, device=dev), pe=torch.zeros(rp, dim, dtype=bf, device=dev),
                  hid=torch.zeros(rp, E, dtype=bf, device=dev), demb=torch.zeros(rp, E, dtype=bf, device=dev),
                  dpre=torch.zeros(rp, E, dtype=bf, device=dev), semb=torch.zeros(rp, E, dtype=bf, device=dev),
                  bias=torch.zeros(total, dtype=torch.float32, device=dev),
                  dsemb=torch.zeros(rows, E, dtype=torch.float32, device=dev))
        eng._arena[key] = ws
    # per-block projections: proj = SiLU(emb) Wt^T + bt
    ws['bias'].zero_()
    K.cast_bf16_colsum(eng.d_emb, ws['dproj'], ws['bias'], rows, total)
    ws['semb'][:rows].copy_(e['semb'])
    lins = e['linears']
    w0, b0 = G.offset[id(lins[0].weight)], G.offset[id(lins[0].bias)]
    offs = [0]
    for lin in lins:
        offs.append(offs[-1] + lin.out_features)
    if all(G.offset[id(l.weight)] == w0 + o * E and G.offset[id(l.bias)] == b0 + o for l, o in zip(lins, offs)):
        # the gradient layout keeps the projections' weights (and biases) contiguous in `linears` order
        # (_record_params): dW_all [total, E] = dproj^T semb in one GEMM, the bias gradients in one copy
        G.flat[b0:b0 + total].copy_(ws['bias'])
        K.gemm_batched((ws['dproj'], rp, total, dict(mn_major=True)), (ws['semb'], rp, E, dict(mn_major=True)),
                       G.flat[w0:w0 + total * E].view(total, E), total, E, rp, out_ld=E)
    else:   # caller-defined layout (gradients requested without a tape-derived order): one GEMM per projection
        for lin, off in zip(lins, offs):
            n = lin.out_features
            G(lin.bias).copy_(ws['bias'][off:off + n])
            K.gemm_batched((ws['dproj'], rp, total, dict(col_base=off, mn_major=True)),
                           (ws['semb'], rp, E, dict(mn_major=True)), G(lin.weight), n, E, rp, out_ld=E)
    K.gemm_batched((ws['dproj'], rp, total), (e['w'], total, E, dict(mn_major=True)), ws['dsemb'], rows, E, total,
                   out_ld=E)
    # time MLP (+ class embedding)
    lin1, lin2, ce = e['lin1'], e['lin2'], e['class_embed']
    pos = e['pos_emb']
    K.time_embed_bwd(e['t'], e['freqs'], pos.dim, E, bool(getattr(pos, 'cos_first', False)), lin1.weight, lin1.bias,
                     lin2.weight, e['emb'], ws['dsemb'], e['y'], ws['pe'], ws['hid'], ws['demb'], ws['dpre'],
                     G(lin1.bias), G(lin2.bias), None if ce is None else G(ce.weight))
    K.gemm_batched((ws['demb'], rp, E, dict(mn_major=True)), (ws['hid'], rp, E, dict(mn_major=True)), G(lin2.weight),
                   E, E, rp, out_ld=E)
    K.gemm_batched((ws['dpre'], rp, E, dict(mn_major=True)), (ws['pe'], rp, pos.dim, dict(mn_major=True)),
                   G(lin1.weight), E, pos.dim, rp, out_ld=pos.dim)


_BWD = {'res': _res_bwd, 'attn': _attn_bwd, 'down': _down_bwd, 'up': _up_bwd, 'first': _first_bwd, 'embed': _embed_bwd}


def run_backward(eng: Engine, tape, dout, params):
    """Replays `tape` in reverse.  dout: fp32 NCHW gradient of the network output."""
    if not dout.is_cuda or dout.dtype != torch.float32:
        raise RuntimeError('backward: expected a float32 CUDA gradient')
    G = _Grads(eng, params, tape)
    eng.flat_grad = G.flat        # the parameters' gradients are views into this buffer, in completion order
    # data-parallel training through b200diff.train.TrainStep: the gradient all-reduce runs INSIDE the backward, bucket by
    # bucket (`eng.ddp_overlap_bytes` > 0 is set by TrainStep for single-micro-batch steps on > 1 ranks; plain autograd /
    # DistributedDataParallel users never get here, their wrapper reduces the returned gradients itself)
    ar = None
    bucket = int(getattr(eng, 'ddp_overlap_bytes', 0) or 0)
    eng.grads_reduced = False
    if bucket > 0 and not G.aliased:
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            ar = _OverlappedAllReduce(G, tape, bucket)
    with torch.no_grad():
        embed_rec = None
        for e in reversed(tape):
            kind = e['kind']
            if kind == 'head':
                _head_bwd(eng, e, G, dout)
            elif kind == 'embed':
                embed_rec = e        # recorded first, but its gradient is complete only after every block ran
                continue
            else:
                _BWD[kind](eng, e, G)
            if ar is not None:
                ar.after_record(e)
        if embed_rec is not None:
            _embed_bwd(eng, embed_rec, G)
        if ar is not None:
            ar.finish()
            eng.grads_reduced = True
            eng.ddp_buckets = ar.n_buckets
    return G.as_tuple()
