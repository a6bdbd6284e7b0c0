"""Forward executor of the UNets: turns one `model(x, t, y)` call into a sequence of libb200diff kernel launches.

Data layout in HBM:
  * residual stream / skip connections: fp32 NHWC  [B, H, W, C]   (written by conv epilogues)
  * tensor-core operands:               bf16 NHWC  [B, H, W, C]   (written once by GroupNorm+SiLU / casts)
  * conv weights: bf16 [Cout, taps*Cin (+Cin_shortcut)] K-major, prepacked from the fp32 nn.Parameters and
    re-packed automatically when a parameter changes (optimizer step, load_state_dict, EMA swap)
  * model input / output: the reference's fp32 NCHW.
All buffers are torch tensors owned by this object (a static arena per batch shape, which is also what makes
the whole forward CUDA-graph capturable); the C library never allocates.

Reference call sites replaced: models/unet.py:121-152 (UNet.forward), :30-43 (ResBlock.forward),
models/modules.py:89-102 (SelfAttentionBlock.forward), models/unet_categorial_adagn.py:44-62,165-208.
"""
import contextlib
from typing import Dict, List, Optional

import torch
import torch.nn as nn

import b200diff as K


class Act:
    """fp32 NHWC activation [B, H, W, C]."""
    __slots__ = ('t', 'B', 'H', 'W', 'C', 'stats', 'g', 'pre')

    def __init__(self, t, B, H, W, C, stats=None):
        self.t, self.B, self.H, self.W, self.C = t, B, H, W, C
        # GN(+SiLU) of this tensor already applied by its producer (Engine.conv_block_out):
        # (GroupNorm module, silu, bf16 NHWC tensor [, bf16 raw copy]); with a raw copy both tensors are
        # [B, H, W, C + C_skip] wide and hold this tensor's part in their first C channels (consumer concatenates a skip)
        self.pre = None
        self.stats = stats   # [B, C, 2] int64 fixed-point per-(image, channel) sum / sum of squares from the producing kernel
        self.g = None        # fp32 NHWC gradient buffer, created by the first backward contribution (models/backward.py)


class Engine:
    # conv1 outputs of ResBlocks (consumed only by the following GroupNorm) are stored as bf16; their GroupNorm
    # statistics still come from the fp32 accumulators in the conv epilogue
    # (inference only; measured: eps rel-L2 +0.6e-3, DDIM-50 CIFAR-10 +2 %, pesser-256 forward +3.4 %)
    h_bf16 = bool(int(__import__('os').environ.get('B200_H_BF16', '1')))
    # conv1 -> norm2 of a ResBlock through the fused b200_conv2d_gn_fwd entry where a conv tile holds whole images
    # (16x16 / 8x8 / 4x4 levels, inference only): the intermediate h is never written and one GroupNorm launch per
    # ResBlock disappears (measured: DDIM-50 CIFAR-10 1035 -> 1057 images/s).  B200_FUSE_GN2=0: two launches (A/B).
    fuse_gn2 = bool(int(__import__('os').environ.get('B200_FUSE_GN2', '1')))
    # whole self-attention block (GroupNorm, q / k / v projections, softmax(q k^T) v, output projection, residual) as ONE
    # launch for the CIFAR-10 UNet's 16x16 blocks (T = 256, C = 256, one head; inference only): q, k, v and the attention
    # output never touch HBM.  B200_ATTN_BLOCK=0: the five-launch form (A/B).
    attn_block = bool(int(__import__('os').environ.get('B200_ATTN_BLOCK', '1')))
    attn_bwd_fused = bool(int(__import__('os').environ.get('B200_ATTN_BWD_FUSED', '1')))   # =0: batched-GEMM adjoint (A/B)
    # a block's last conv also emits the GroupNorm(+SiLU) of its output for the NEXT block's norm1 / the output head
    # (b200_conv2d_gn_fwd, block-output form) when that norm has this tensor as its only input; =0: separate launch (A/B)
    fuse_gn1 = bool(int(__import__('os').environ.get('B200_FUSE_GN1', '1')))
    # ... and when the consumer concatenates a skip connection: the producer writes its part of the normalised and raw
    # concat operands, the GroupNorm launch of the consumer handles the skip's channels only; =0: full launch (A/B)
    fuse_gn1_cat = bool(int(__import__('os').environ.get('B200_FUSE_GN1_CAT', '1')))
    # the block-output form also in training forwards (the fp32 output + statistics are written as before: the adjoint of
    # the consumer's GroupNorm reads them; its normalised / raw bf16 operands get per-block buffers); =0: inference only (A/B)
    fuse_gn1_train = bool(int(__import__('os').environ.get('B200_FUSE_GN1_TRAIN', '1')))
    # ... and when nobody reads the fp32 form of such a block output (decoder blocks feeding a concatenating block or the
    # output head: they read the producer-applied bf16 tensors only) it is not written at all; =0: always written (A/B)
    skip_dead_out = bool(int(__import__('os').environ.get('B200_SKIP_DEAD_OUT', '1')))
    # up-sampling conv outputs whose only consumer is a concatenating block's GroupNorm are stored as bf16 (=0: fp32, A/B)
    up_bf16 = bool(int(__import__('os').environ.get('B200_UP_BF16', '1')))
    # first convolution on the tensor cores (64-channel hi / lo / hi split pixels + the first block's GroupNorm fused).
    # Off by default: same-box A/B (DDIM-50, batch 256) 1205.6 / 1202.5 images/s with it, 1205.8 / 1208.4 without -- the
    # 32x32 multi-tile fused epilogue costs what the FP32-FMA kernel + its GroupNorm launch cost.  =1 enables it (A/B).
    first_tc = bool(int(__import__('os').environ.get('B200_FIRST_TC', '0')))

    def __init__(self, model: nn.Module):
        self.model = model
        # 'bf16': operands rounded to bf16 once, fp32 accumulation (eps rel-L2 ~6e-3 vs the fp32 reference);
        # 'fp32': every GEMM operand as a bf16 hi/lo pair, 3 tensor-core terms per product (csrc/precise.cu), exact SiLU /
        #         softmax, fp64 GroupNorm statistics: eps rel-L2 <= 1e-4 (north_star's FP32 gate), ~3x the MMA work.
        self.precision = __import__('os').environ.get('B200_PRECISION', 'bf16')
        self._packed: Dict = {}
        self._const: Dict = {}
        self._sig = None
        self._arena: Dict = {}
        self._stats: Dict = {}
        self._stats_pools: Dict = {}     # input shape -> [[pool tensor, elements used]]
        self._cur_key = None             # input shape (+ lane) of the forward being issued
        self.arena_reuse = __import__('os').environ.get('B200_ARENA_REUSE', '1') != '0'
        self._scratch: Dict = {}         # (input shape, lane) -> scratch chunks of block temporaries (see buf / scope)
        self._scope_depth = 0
        self.pingpong = False            # set by the models from the bottleneck on (inference): see buf()
        self._pp_count: Dict = {}
        self.lane = 0                    # buffer-set index: forwards issued concurrently on different streams (the sampling
                                         # runner's two half-batch branches) use different lanes = disjoint buffers
        self._device = None          # cached per forward (refresh re-reads it)
        self._pt: Dict = {}          # table-managed packed weights: key -> (result, [b200_pack_entry bytes])
        self._pt_table = None        # device table over all entries (rebuilt when an entry is added)
        self._pt_scratch: List = []  # single-entry tables of first-use packs (kept alive until the stream consumed them)
        self.tape = None         # training forward: list of records replayed in reverse by models/backward.py
        self.seed_dev = None     # int64 [1] device scalar: base seed of this forward's dropout masks (graph replayable)
        self.emb_override = None  # [1, total] fp32: precomputed embedding projections of the current timestep (sampling runner)
        self._embed_args = None   # modules of the last time-only embedding (what embed_rows re-evaluates for a whole schedule)
        self._n_drop = 0
        self._epoch = 0           # bumped by invalidate(): part of the weight signature
        self._fwd_gen: Dict = {}  # input shape -> number of forwards that (re)wrote the arena buffers of that shape

    # ------------------------------------------------------------------------------------------
    # precision mode
    # ------------------------------------------------------------------------------------------
    @property
    def split(self) -> bool:
        return self.precision == 'fp32'

    @property
    def m3(self) -> int:
        """Channel multiplier of GEMM operands: 3 in FP32 mode ([hi | lo | hi] along the contraction), else 1."""
        return 3 if self.precision == 'fp32' else 1

    def set_precision(self, precision: str):
        """'bf16' (default) or 'fp32' (see __init__).  Inference only in 'fp32'; own UNet families (models/unet.py,
        models/unet_categorial_adagn.py).  Switching drops the packed operands and the captured sampling graphs."""
        if precision not in ('bf16', 'fp32'):
            raise ValueError(f"precision must be 'bf16' or 'fp32', got {precision!r}")
        if precision != self.precision:
            self.precision = precision
            self.invalidate()

    def _fp32_inference_only(self, what):
        if self.split and self.tape is not None:
            raise RuntimeError(f"{what}: precision='fp32' is an inference mode (the training step runs in bf16 operands)")

    # ------------------------------------------------------------------------------------------
    # buffers and packed weights
    # ------------------------------------------------------------------------------------------
    @property
    def device(self):
        d = self._device
        if d is None:
            d = self._device = next(self.model.parameters()).device
        return d

    def buf(self, tag, shape, dtype, temp=False):
        """Arena buffer `tag` of the current forward.  Training forwards (tape) keep one buffer per tag: the backward reads
        them.  Inference forwards reuse memory by liveness (B200_ARENA_REUSE=0 disables):
          * temp=True inside a `scope()` (a block's bf16 operands, intermediates, attention workspaces): bump-allocated
            from a scratch pool and released when the block's scope ends;
          * block outputs while `pingpong` is set (bottleneck + decoder: nothing there is a skip connection): every block
            reads only its predecessor's output, so outputs of equal shape alternate between two buffers.
        Encoder outputs (the skips), embeddings and the statistics keep their own buffers.  The request sequence of a
        forward is deterministic, so every buffer gets the same address in every forward (CUDA-graph replayable)."""
        if self.tape is None and self.arena_reuse:
            if temp and self._scope_depth > 0:
                return self._temp(tag, tuple(shape), dtype)
            if self.pingpong and isinstance(tag, str) and tag.endswith('.out'):
                k = (tuple(shape), dtype)
                n = self._pp_count.get(k, 0)
                self._pp_count[k] = n + 1
                tag = f'pingpong{n & 1}.out'
        key = (tag, tuple(shape), dtype, self.device, self.lane)
        t = self._arena.get(key)
        if t is None:
            t = torch.empty(shape, dtype=dtype, device=self.device)
            self._arena[key] = t
        return t

    def _temp(self, tag, shape, dtype):
        """Bump allocation from the scratch chunks of the current (input shape, lane); 1 KB granularity (TMA operands need
        128-byte alignment).  Chunks grow geometrically during the first forward of a shape and are then stable."""
        n = 1
        for d in shape:
            n *= d
        nbytes = n * torch.empty((), dtype=dtype).element_size()
        need = (nbytes + 1023) // 1024 * 1024
        st = self._scratch.get(self._cur_key)
        if st is None:
            st = self._scratch[self._cur_key] = {'chunks': [], 'pos': [0, 0], 'views': {}, 'peak': 0}
        chunks, pos = st['chunks'], st['pos']
        while not (pos[0] < len(chunks) and pos[1] + need <= chunks[pos[0]].numel()):
            if pos[0] < len(chunks):
                pos[0] += 1
                pos[1] = 0
            else:
                total = sum(c.numel() for c in chunks)
                chunks.append(torch.empty(max(need, total, 64 << 20), dtype=torch.uint8, device=self.device))
                pos[1] = 0
        vkey = (tag, shape, dtype, pos[0], pos[1])
        v = st['views'].get(vkey)
        if v is None:
            v = st['views'][vkey] = chunks[pos[0]][pos[1]:pos[1] + nbytes].view(dtype).view(shape)
        pos[1] += need
        return v

    @contextlib.contextmanager
    def scope(self):
        """Lifetime of a block's temporaries: buffers requested with temp=True inside are released at exit."""
        st = self._scratch.get(self._cur_key)
        saved = list(st['pos']) if st is not None else [0, 0]
        self._scope_depth += 1
        try:
            yield
        finally:
            self._scope_depth -= 1
            st = self._scratch.get(self._cur_key)
            if st is not None:
                st['pos'][0], st['pos'][1] = saved

    def arena_bytes(self) -> int:
        """Device memory held by activation buffers (arena + scratch chunks), all shapes and lanes."""
        n = sum(t.numel() * t.element_size() for t in self._arena.values())
        return n + sum(c.numel() for st in self._scratch.values() for c in st['chunks'])

    def stats_buf(self, tag, B, C):
        """[B, C, 2] int64 fixed-point accumulator (K.STAT_Q1 / K.STAT_Q2) for the GroupNorm statistics of a conv output
        (zeroed every forward).  Integer atomics: order-independent, so forwards are bitwise reproducible.
        Buffers and their pools belong to the input shape of the current forward (`check_input`), like the activation
        arena: a forward of another shape neither reuses nor zeroes them (a training tape may still need them)."""
        fk = self._cur_key
        key = (tag, B, C, fk, self.device)
        t = self._stats.get(key)
        if t is None:
            # sub-allocated from a few large pool tensors so that one forward needs one memset per pool, not one
            # tiny kernel per GroupNorm input (70 of them for the CIFAR-10 UNet)
            n = B * C * 2
            pools = self._stats_pools.setdefault(fk, [])
            if not pools or pools[-1][1] + n > pools[-1][0].numel() or pools[-1][0].device != self.device:
                cap = max(n, 1 << 20)
                pools.append([torch.zeros(cap, dtype=torch.int64, device=self.device), 0])
            pool = pools[-1]
            t = pool[0][pool[1]:pool[1] + n].view(B, C, 2)
            pool[1] += (n + 63) // 64 * 64
            self._stats[key] = t
        return t

    def begin_forward(self):
        self.refresh()

    def refresh(self):
        """Brings the packed bf16 weights up to date when a parameter was modified (optimizer step, load_state_dict,
        EMA swap): table-managed packs are re-created in place by ONE b200_pack_weights launch; everything is dropped
        when a parameter moved (new storage)."""
        self._device = None
        sig = (self._epoch,) + tuple((p.data_ptr(), p._version) for p in self.model.parameters())
        if sig != self._sig:
            moved = self._sig is None or len(sig) != len(self._sig) or sig[0] != self._sig[0] or \
                any(a[0] != b[0] for a, b in zip(sig[1:], self._sig[1:]))
            self._packed.clear()
            if moved:
                self._pt.clear()
                self._pt_table = None
            elif self._pt:
                if self._pt_table is None:
                    blob = b''.join(e for _, entries in self._pt.values() for e in entries)
                    n = sum(len(entries) for _, entries in self._pt.values())
                    self._pt_table = (torch.frombuffer(bytearray(blob), dtype=torch.uint8).to(self.device), n)
                K.pack_weights(*self._pt_table)
            self._pt_scratch.clear()
            self._sig = sig
        dev = self.device
        if dev.type != 'cuda':
            raise RuntimeError('b200diff models run on a CUDA device only: there is no CPU/PyTorch fallback '
                               f'(parameters are on {dev})')

    def invalidate(self):
        """Forces a full re-pack of the bf16 operands (and re-capture of the sampling graphs) at the next forward.
        Needed only after writes that bypass the parameters' version counters -- `p.data.copy_(...)`, `p.data = ...` on
        the same storage, raw-pointer writes from another library; `p.copy_()` under no_grad, optimizers,
        `load_state_dict` and models.EMA are detected automatically through (data_ptr, _version)."""
        self._epoch += 1

    def const(self, key, make):
        """Parameter-independent device constants (frequency tables): survive weight updates, unlike `packed`."""
        v = self._const.get(key)
        if v is None:
            v = self._const[key] = make()
        return v

    def packed(self, key, make):
        v = self._packed.get(key)
        if v is None:
            with torch.no_grad():
                v = make()
            self._packed[key] = v
        return v

    def pack_table(self, key, result, entries):
        """Registers a table-managed pack (`entries`: b200_pack_entry byte strings writing into the tensors of `result`)
        and fills it now; later parameter updates re-run all registered entries in one launch (refresh)."""
        self._pt[key] = (result, entries)
        self._pt_table = None
        blob = torch.frombuffer(bytearray(b''.join(entries)), dtype=torch.uint8).to(self.device)
        self._pt_scratch.append(blob)
        K.pack_weights(blob, len(entries))
        return result

    def w_conv(self, tag, conv: nn.Conv2d, shortcut: Optional[nn.Conv2d] = None):
        """bf16 [Cout, taps*Cin (+ Cin_shortcut)] forward operand + fp32 bias (conv bias, summed with the fused shortcut's).
        FP32 mode: [Cout, 3*(taps*Cin + Cin_shortcut)], per tap [w_hi | w_hi | w_lo] (pack mode 3)."""
        key = ('conv', tag)
        hit = self._pt.get(key)
        if hit is not None:
            return hit[0]
        Co, Ci = conv.weight.shape[:2]
        taps = conv.weight[0, 0].numel()
        m, mode = self.m3, (3 if self.split else 0)
        Kd = m * (taps * Ci + (shortcut.in_channels if shortcut is not None else 0))
        w = torch.empty((Co, Kd), dtype=torch.bfloat16, device=self.device)
        entries = [K.pack_entry_bytes(conv.weight, w, Co, Ci, taps, mode, ld=Kd)]
        b = conv.bias
        if shortcut is not None:
            entries.append(K.pack_entry_bytes(shortcut.weight, w, Co, shortcut.in_channels, 1, mode, col0=m * taps * Ci, ld=Kd))
            b = torch.empty(Co, dtype=torch.float32, device=self.device)
            entries.append(K.pack_entry_bytes(conv.bias, b, Co, 1, 1, 2, src2=shortcut.bias))
        return self.pack_table(key, (w, b), entries)

    def w_dgrad(self, tag, conv: nn.Conv2d):
        """bf16 [Cin, taps*Cout] data-gradient operand: spatially flipped taps, in/out channels swapped."""
        key = ('dgrad', tag)
        hit = self._pt.get(key)
        if hit is not None:
            return hit[0]
        Co, Ci = conv.weight.shape[:2]          # Conv2d [Co, Ci, kh, kw] or Conv1d [Co, Ci, 1]
        taps = conv.weight[0, 0].numel()
        w = torch.empty((Ci, taps * Co), dtype=torch.bfloat16, device=self.device)
        return self.pack_table(key, w, [K.pack_entry_bytes(conv.weight, w, Co, Ci, taps, 1, ld=taps * Co)])

    def w_up2(self, tag, conv: nn.Conv2d):
        return self.packed(('up2', tag), lambda: (K.pack_weight_up2(conv.weight, split=self.split),
                                                  conv.bias.detach().float().contiguous()))

    def w_tproj(self, linears: List[nn.Linear]):
        """All per-ResBlock embedding projections stacked into one [sum(out), E] bf16 matrix."""
        key = ('tproj',)
        hit = self._pt.get(key)
        if hit is not None:
            return hit[0]
        E = linears[0].in_features
        total = sum(l.out_features for l in linears)
        m, mode = self.m3, (3 if self.split else 0)
        w = torch.empty((total, m * E), dtype=torch.bfloat16, device=self.device)
        b = torch.empty(total, dtype=torch.float32, device=self.device)
        entries, off = [], 0
        for l in linears:
            n = l.out_features
            entries.append(K.pack_entry_bytes(l.weight, w, n, E, 1, mode, row0=off, ld=m * E))
            entries.append(K.pack_entry_bytes(l.bias, b[off:off + n], n, 1, 1, 2))
            off += n
        return self.pack_table(key, (w, b), entries)

    # ------------------------------------------------------------------------------------------
    # ops
    # ------------------------------------------------------------------------------------------
    def next_drop_seed(self):
        """Per-block offset added (mod 2^64, on the device) to the per-forward base seed in `seed_dev`."""
        self._n_drop += 1
        return (0x9E3779B97F4A7C15 * self._n_drop) & 0xFFFFFFFFFFFFFFFF

    def new_dropout_base(self):
        """Draws the per-forward base seed ON THE DEVICE (torch's CUDA generator: reproducible under manual_seed and
        re-drawn on every replay of a captured graph)."""
        if self.seed_dev is None or self.seed_dev.device != self.device:
            self.seed_dev = torch.zeros(1, dtype=torch.int64, device=self.device)
        self.seed_dev.random_()
        self._n_drop = 0

    def dropout_seed_of(self, rec) -> int:
        """Effective 64-bit seed of a tape record's dropout mask (host sync; for tests that regenerate the mask)."""
        return (rec['drop_seed'] + int(self.seed_dev.item())) & 0xFFFFFFFFFFFFFFFF

    def gn(self, tag, x: Act, skip: Optional[Act], norm: nn.GroupNorm, silu=True, raw=False, scale=None, shift=None,
           ss_ld=0, resample=0, drop_p=0.0, drop_seed=0):
        if (x.pre is not None and x.pre[0] is norm and x.pre[1] == bool(silu) and scale is None and resample == 0
                and drop_p == 0.0 and not self.split):
            if len(x.pre) == 3 and skip is None and not raw:
                return x.pre[2], None      # the producing conv already applied this GroupNorm (fuse_gn1)
            if (len(x.pre) == 4 and skip is not None and skip.stats is not None
                    and x.pre[2].shape[-1] == x.C + skip.C):
                # the producer wrote x's part of the normalised and raw concat operands: only the skip's part is left
                K.groupnorm_apply(None, x.C, None, skip.t, skip.C, skip.stats, x.B, x.H * x.W, x.W, norm.num_groups,
                                  norm.weight, norm.bias, norm.eps, x.pre[2], silu=silu, raw_out=x.pre[3])
                return x.pre[2], (x.pre[3] if raw else None)
        seed_dev = self.seed_dev if drop_p > 0 else None
        C = x.C + (skip.C if skip is not None else 0)
        Ho, Wo = x.H, x.W
        if resample == 1:
            Ho, Wo = x.H // 2, x.W // 2
        elif resample == 2:
            Ho, Wo = x.H * 2, x.W * 2
        if self.split:
            self._fp32_inference_only(tag)
            if x.stats is None or (skip is not None and skip.stats is None):
                raise RuntimeError(f"{tag}: precision='fp32' needs producer statistics for every GroupNorm input "
                                   '(parameter-free resampling layers are not supported in this mode)')
            out = self.buf(tag + '.gn3', (x.B, Ho, Wo, 3 * C), torch.bfloat16, temp=True)
            raw_out = self.buf(tag + '.raw3', (x.B, x.H, x.W, 3 * C), torch.bfloat16, temp=True) if raw else None
            K.groupnorm_apply_split(x.t, x.C, x.stats, None if skip is None else skip.t, 0 if skip is None else skip.C,
                                    None if skip is None else skip.stats, x.B, x.H * x.W, x.W, norm.num_groups,
                                    norm.weight, norm.bias, norm.eps, out, scale=scale, shift=shift, ss_ld=ss_ld,
                                    silu=silu, resample=resample, raw_out=raw_out)
            return out, raw_out
        out = self.buf(tag + '.gn', (x.B, Ho, Wo, C), torch.bfloat16, temp=True)
        raw_out = self.buf(tag + '.raw', (x.B, x.H, x.W, C), torch.bfloat16, temp=True) if raw else None
        if x.stats is not None and (skip is None or skip.stats is not None):
            K.groupnorm_apply(x.t, x.C, x.stats, None if skip is None else skip.t, 0 if skip is None else skip.C,
                              None if skip is None else skip.stats, x.B, x.H * x.W, x.W, norm.num_groups,
                              norm.weight, norm.bias, norm.eps, out, scale=scale, shift=shift, ss_ld=ss_ld, silu=silu,
                              resample=resample, raw_out=raw_out, drop_p=drop_p, drop_seed=drop_seed,
                              drop_seed_dev=seed_dev)
        else:
            if drop_p > 0 or self.tape is not None:
                raise RuntimeError(f'{tag}: training needs producer statistics for every GroupNorm input '
                                   '(parameter-free resampling layers are inference-only)')
            K.groupnorm_silu(x.t, x.C, None if skip is None else skip.t, 0 if skip is None else skip.C, x.B,
                             x.H * x.W, x.W, norm.num_groups, norm.weight, norm.bias, norm.eps, out, scale=scale,
                             shift=shift, ss_ld=ss_ld, silu=silu, resample=resample, raw_out=raw_out)
        return out, raw_out

    def conv3x3(self, tag, a, B, H, W, Cin, conv, *, rowadd=None, rowadd_ld=0, residual: Optional[Act] = None,
                sc_a=None, sc_C=0, sc_conv=None, out_mode=K.OUT_F32_NHWC, out=None, intermediate=False, temp=False):
        Cout = conv.out_channels
        w, b = self.w_conv(tag, conv, sc_conv)
        stats = None
        m = self.m3      # FP32 mode: `a` / `sc_a` hold [hi | lo | hi] per pixel, the packed weights [hi | hi | lo] per tap
        if out is None:
            if intermediate and self.h_bf16 and Cout >= 128 and self.tape is None and not self.split:   # narrow toy nets keep fp32 h
                out_mode = K.OUT_BF16_NHWC
            out = self.buf(tag + '.out', (B, H, W, Cout),
                           torch.bfloat16 if out_mode == K.OUT_BF16_NHWC else torch.float32, temp=temp or intermediate)
            stats = self.stats_buf(tag, B, Cout) if Cout > 32 else None
        K.conv2d(a, w, Cout, B, H, W, K.taps_3x3_s1(), a0_geom=(m * Cin, H, W, 1),
                 a1=sc_a, a1_geom=(m * sc_C, H, W, 1) if sc_a is not None else None, bias=b,
                 alg_macs=float(B) * H * W * Cout * (9 * Cin + sc_C),
                 rowadd=rowadd, rowadd_ld=rowadd_ld, residual=None if residual is None else residual.t,
                 res_ld=0 if residual is None else residual.C, out=out, out_mode=out_mode, stats=stats)
        return Act(out, B, H, W, Cout, stats)

    def attention(self, tag, blk, x: Act, next_gn=None) -> Act:
        """models/modules.py:89-102 (own UNets): separate q, k, v, proj 1x1 convs; q scaled by d^-1/2.
        next_gn: see `resblock` (applied by the output projection's epilogue on the five-launch path)."""
        if self.split:
            return self._attention_split(tag, blk, x)
        if (self.attn_block and self.tape is None and x.stats is not None
                and K.attn_block_ok(x.H * x.W, x.C, blk.n_heads, blk.norm.num_groups)):
            return self._attention_block_fused(tag, blk, x)
        key = ('attn', tag)
        hit = self._pt.get(key)
        if hit is None:
            C = blk.q.out_channels
            bf, dev = torch.bfloat16, self.device
            wqk, wv, wp = (torch.empty((2 * C, C), dtype=bf, device=dev), torch.empty((C, C), dtype=bf, device=dev),
                           torch.empty((C, C), dtype=bf, device=dev))
            bqk = torch.empty(2 * C, dtype=torch.float32, device=dev)
            entries = [K.pack_entry_bytes(blk.q.weight, wqk, C, C, 1, 0, ld=C),
                       K.pack_entry_bytes(blk.k.weight, wqk, C, C, 1, 0, row0=C, ld=C),
                       K.pack_entry_bytes(blk.q.bias, bqk, C, 1, 1, 2), K.pack_entry_bytes(blk.k.bias, bqk[C:], C, 1, 1, 2),
                       K.pack_entry_bytes(blk.v.weight, wv, C, C, 1, 0, ld=C),
                       K.pack_entry_bytes(blk.proj.weight, wp, C, C, 1, 0, ld=C)]
            weights = self.pack_table(key, (wqk, bqk, wv, blk.v.bias, wp, blk.proj.bias), entries)
        else:
            weights = hit[0]
        return self.attention_core(tag, x, blk.norm, weights, blk.n_heads, blk.scale,
                                   mods=(blk.q, blk.k, blk.v, blk.proj), next_gn=next_gn)

    def _attention_block_fused(self, tag, blk, x: Act) -> Act:
        """models/modules.py:89-102 in one launch (b200_attn_block_fwd).  The caller checked K.attn_block_ok."""
        key = ('attnblk', tag)
        hit = self._pt.get(key)
        if hit is None:
            C = blk.q.out_channels
            w = torch.empty((4 * C, C), dtype=torch.bfloat16, device=self.device)
            b = torch.empty(4 * C, dtype=torch.float32, device=self.device)
            entries = []
            for i, mod in enumerate((blk.q, blk.k, blk.v, blk.proj)):
                entries.append(K.pack_entry_bytes(mod.weight, w, C, C, 1, 0, row0=i * C, ld=C))
                entries.append(K.pack_entry_bytes(mod.bias, b[i * C:], C, 1, 1, 2))
            w, b = self.pack_table(key, (w, b), entries)
        else:
            w, b = hit[0]
        B, H, W, C = x.B, x.H, x.W, x.C
        out = self.buf(tag + '.out', (B, H, W, C), torch.float32)
        stats = self.stats_buf(tag, B, C)
        K.attn_block(x.t, x.stats, blk.norm.weight, blk.norm.bias, blk.norm.eps, w, b, out, stats, B, H * W, C,
                     blk.n_heads, blk.norm.num_groups, blk.scale)
        return Act(out, B, H, W, C, stats)

    def _attention_split(self, *args, **kwargs):
        with self.scope():       # the block's temporaries are released when it returns
            return self._attention_split_body(*args, **kwargs)

    def _attention_split_body(self, tag, blk, x: Act) -> Act:
        """FP32-mode attention block (models/modules.py:89-102): GroupNorm -> q, k, v 1x1 convs with fp32 outputs ->
        S = q k^T and O = softmax(S d^-1/2) v as batched tensor-core GEMMs over split operands (K = 3d resp. 3T), exact
        row softmax in between -> 1x1 proj + residual.  [B*h, T, T] scores are materialised (accuracy mode)."""
        self._fp32_inference_only(tag)
        B, H, W, C = x.B, x.H, x.W, x.C
        T, heads = H * W, blk.n_heads
        d = C // heads
        if T % 8 != 0 or d % 8 != 0 or (heads > 1 and (3 * d) % 64 != 0):
            raise RuntimeError(f"attention block {tag}: T={T}, head_dim={d} not supported in precision='fp32'")
        key = ('attn3', tag)
        hit = self._pt.get(key)
        if hit is None:
            bf, dev = torch.bfloat16, self.device
            wqkv = torch.empty((3 * C, 3 * C), dtype=bf, device=dev)
            wp = torch.empty((C, 3 * C), dtype=bf, device=dev)
            bqkv = torch.empty(3 * C, dtype=torch.float32, device=dev)
            entries = []
            for i, mod in enumerate((blk.q, blk.k, blk.v)):
                entries.append(K.pack_entry_bytes(mod.weight, wqkv, C, C, 1, 3, row0=i * C, ld=3 * C))
                entries.append(K.pack_entry_bytes(mod.bias, bqkv[i * C:], C, 1, 1, 2))
            entries.append(K.pack_entry_bytes(blk.proj.weight, wp, C, C, 1, 3, ld=3 * C))
            weights = self.pack_table(key, (wqkv, bqkv, wp, blk.proj.bias), entries)
        else:
            weights = hit[0]
        wqkv, bqkv, wp, bp = weights
        n, _ = self.gn(tag, x, None, blk.norm, silu=False)                       # [B, T, 3C] split
        qkv = self.buf(tag + '.qkv32', (B, T, 3 * C), torch.float32, temp=True)
        K.conv2d(n, wqkv, 3 * C, B, H, W, K.taps_1x1(), a0_geom=(3 * C, H, W, 1), bias=bqkv, out=qkv,
                 alg_macs=float(B) * T * 3 * C * C)
        bf = torch.bfloat16
        qs = self.buf(tag + '.q3', (B, T, 3 * C), bf, temp=True)      # per head [q_hi | q_lo | q_hi]
        ks = self.buf(tag + '.k3', (B, T, 3 * C), bf, temp=True)      # per head [k_hi | k_hi | k_lo]
        vs = self.buf(tag + '.v3', (B, 3, T, C), bf, temp=True)       # planes [v_hi; v_hi; v_lo] along the key dimension
        K.split_cast(qkv, qs, B * T, C, in_ld=3 * C, in_col0=0, group=d, pattern=K.SPLIT_ACT)
        K.split_cast(qkv, ks, B * T, C, in_ld=3 * C, in_col0=C, group=d, pattern=K.SPLIT_WEIGHT)
        K.split_cast(qkv, vs, B * T, C, in_ld=3 * C, in_col0=2 * C, pattern=K.SPLIT_WEIGHT, planes_rows=T)
        G_ = B * heads
        S = self.buf('attn3_ws.S', (G_, T, T), torch.float32, temp=True)
        P = self.buf('attn3_ws.P', (G_, T, 3 * T), bf, temp=True)
        grid = dict(batch=B, heads=heads)
        K.gemm_batched((qs, T, 3 * C, dict(col_base=0, col_head=3 * d)), (ks, T, 3 * C, dict(col_base=0, col_head=3 * d)), S,
                       T, T, 3 * d, **grid, out_ld=T, out_batch_stride=heads * T * T, out_head_stride=T * T)
        K.softmax_rows_split(S, P, G_ * T, T, blk.scale)
        o = self.buf(tag + '.o32', (B, T, C), torch.float32, temp=True)
        K.gemm_batched((P, T, 3 * T, dict(per_head_batch=True)), (vs, 3 * T, C, dict(col_head=d, mn_major=True)), o,
                       T, d, 3 * T, **grid, out_ld=C, out_batch_stride=T * C, out_head_stride=d)
        os_ = self.buf(tag + '.o3', (B, T, 3 * C), bf, temp=True)
        K.split_cast(o, os_, B * T, C)
        out = self.buf(tag + '.out', (B, H, W, C), torch.float32)
        stats = self.stats_buf(tag, B, C)
        K.conv2d(os_, wp, C, B, H, W, K.taps_1x1(), a0_geom=(3 * C, H, W, 1), bias=bp, residual=x.t, res_ld=C, out=out,
                 stats=stats, alg_macs=float(B) * T * C * C)
        return Act(out, B, H, W, C, stats)

    def attention_core(self, *args, **kwargs):
        with self.scope():       # the block's temporaries are released when it returns
            return self._attention_core_body(*args, **kwargs)

    def _attention_core_body(self, tag, x: Act, norm: nn.GroupNorm, weights, heads: int, scale: float, mods=None,
                             next_gn=None) -> Act:
        """GroupNorm -> [q|k] and v^T 1x1 convs -> fused softmax(q k^T * scale) v -> 1x1 proj + residual.
        `weights` = (Wqk [2C, C] bf16 with rows [q heads..., k heads...], bqk, Wv [C, C], bv, Wproj, bproj)."""
        if self.split:
            raise RuntimeError(f"attention block {tag}: precision='fp32' is implemented for the reference's own UNet "
                               'families (models/unet.py, models/unet_categorial_adagn.py), not for ADM / pesser')
        B, H, W, C = x.B, x.H, x.W, x.C
        T = H * W
        d = C // heads
        if T % 8 != 0 or not ((T <= 256 and d in (64, 128, 256)) or (T <= 256 and d % 64 == 0 and d <= 512 and heads == 1)
                              or (T > 256 and d == 64)):
            raise RuntimeError(f'attention block {tag}: T={T}, head_dim={d}, heads={heads} not supported by '
                               'b200_attention_fwd')
        wqk, bqk, wv, bv, wp, bp = weights
        n, _ = self.gn(tag, x, None, norm, silu=False)
        qk = self.buf(tag + '.qk', (B, T, 2 * C), torch.bfloat16, temp=True)
        K.conv2d(n, wqk, 2 * C, B, H, W, K.taps_1x1(), a0_geom=(C, H, W, 1), bias=bqk, out=qk,
                 out_mode=K.OUT_BF16_NHWC)
        vt = self.buf(tag + '.vt', (B, C, T), torch.bfloat16, temp=True)
        K.conv2d(n, wv, C, B, H, W, K.taps_1x1(), a0_geom=(C, H, W, 1), bias=bv, out=vt, out_mode=K.OUT_BF16_NCHW)
        o = self.buf(tag + '.o', (B, T, C), torch.bfloat16, temp=True)
        # training: keep the per-row log-sum-exp so that the one-launch adjoint (b200_attention_bwd) can recompute P
        lse = (self.buf(tag + '.lse', (B, heads, T), torch.float32, temp=True)
               if self.tape is not None and self.attn_bwd_fused and K.attention_bwd_ok(T, d) else None)
        K.attention(qk, 2 * C, 0, C, vt, o, C, B, T, heads, d, scale, lse=lse)
        if C % 64 == 0 and self.next_gn_ok(B, H, W, C, next_gn):
            res = self.conv_block_out(tag, o, B, H, W, C, C, wp, bp, K.taps_1x1(), next_gn, residual=x)
        else:
            out = self.buf(tag + '.out', (B, H, W, C), torch.float32)
            stats = self.stats_buf(tag, B, C)
            K.conv2d(o, wp, C, B, H, W, K.taps_1x1(), a0_geom=(C, H, W, 1), bias=bp, residual=x.t, res_ld=C, out=out,
                     stats=stats)
            res = Act(out, B, H, W, C, stats)
        if self.tape is not None:
            if mods is None:
                raise RuntimeError(f'{tag}: this attention block did not register its parameters for the backward pass')
            self.tape.append(dict(kind='attn', tag=tag, x=x, out=res, norm=norm, n=n, qk=qk, vt=vt, o=o, lse=lse,
                                  heads=heads, scale=scale, mods=mods))
        return res

    def downsample_conv(self, *args, **kwargs):
        with self.scope():       # the block's temporaries are released when it returns
            return self._downsample_conv_body(*args, **kwargs)

    def _downsample_conv_body(self, tag, conv: nn.Conv2d, x: Act, pad_lo=1) -> Act:
        B, H, W, C = x.B, x.H, x.W, x.C
        m = self.m3
        planes = self.buf(tag + '.planes', (B, 4, H // 2, W // 2, m * C), torch.bfloat16, temp=True)
        if self.split:
            self._fp32_inference_only(tag)
            K.split_cast(x.t, planes, B * H * W, C, parity_hw=(H, W))
        else:
            K.cast_bf16(x.t, planes, B, H, W, C, parity_split=True)
        w, b = self.w_conv(tag, conv)
        Cout = conv.out_channels
        out = self.buf(tag + '.out', (B, H // 2, W // 2, Cout), torch.float32)
        stats = self.stats_buf(tag, B, Cout)
        K.conv2d(planes, w, Cout, B, H // 2, W // 2, K.taps_3x3_s2(pad_lo), a0_geom=(m * C, H // 2, W // 2, 4), bias=b,
                 out=out, stats=stats, alg_macs=9.0 * B * (H // 2) * (W // 2) * C * Cout)
        res = Act(out, B, H // 2, W // 2, Cout, stats)
        if self.tape is not None:
            self.tape.append(dict(kind='down', tag=tag, x=x, out=res, conv=conv, planes=planes, pad_lo=pad_lo))
        return res

    def upsample_conv(self, *args, **kwargs):
        with self.scope():       # the block's temporaries are released when it returns
            return self._upsample_conv_body(*args, **kwargs)

    def _upsample_conv_body(self, tag, conv: nn.Conv2d, x: Act, bf16_out: bool = False) -> Act:
        """nearest-2x + conv3x3 as four 2x2-tap phase convolutions on the low-res grid (2.25x fewer MACs)."""
        B, H, W, C = x.B, x.H, x.W, x.C
        if self.tape is not None:
            # training: materialise the nearest-2x tensor and run the plain 3x3 conv, whose adjoints are the standard
            # dgrad / wgrad kernels (the 4-phase form would need phase-wise weight-gradient folding)
            ub = self.buf(tag + '.up', (B, 2 * H, 2 * W, C), torch.bfloat16)
            K.upsample2_bf16(x.t, ub, B, H, W, C)
            res = self.conv3x3(tag, ub, B, 2 * H, 2 * W, C, conv)
            self.tape.append(dict(kind='up', tag=tag, x=x, out=res, conv=conv, ub=ub))
            return res
        m = self.m3
        xb = self.buf(tag + '.bf16', (B, H, W, m * C), torch.bfloat16, temp=True)
        if self.split:
            K.split_cast(x.t, xb, B * H * W, C)
        else:
            K.cast_bf16(x.t, xb, B, H, W, C)
        w, b = self.w_up2(tag, conv)
        Cout = conv.out_channels
        # bf16_out: the caller guarantees that the ONLY consumer is the GroupNorm (+ fused 1x1 shortcut) of a block that
        # concatenates a skip connection, i.e. nobody reads this tensor as an fp32 residual: like the ResBlock
        # intermediates (h_bf16) it is then stored as bf16 -- its GroupNorm statistics still come from the fp32 accumulators
        b16 = bf16_out and self.up_bf16 and self.h_bf16 and Cout >= 128 and not self.split
        out = self.buf(tag + '.out', (B, 2 * H, 2 * W, Cout), torch.bfloat16 if b16 else torch.float32)
        stats = self.stats_buf(tag, B, Cout)
        K.conv2d(xb, w, Cout, B, H, W, K.taps_up2_3x3(), a0_geom=(m * C, H, W, 1), bias=b, out=out,
                 out_mode=K.OUT_BF16_NHWC if b16 else K.OUT_F32_NHWC,
                 w_rows_per_phase=Cout, stats=stats, alg_macs=9.0 * B * 4 * H * W * C * Cout)
        return Act(out, B, 2 * H, 2 * W, Cout, stats)

    def resample_plain(self, tag, x: Act, mode: int) -> Act:
        """Parameter-free 2x2 average pool (mode 1) / nearest 2x (mode 2) of the fp32 residual stream (the
        conv_resample=False / with_conv=False variants).  The result carries no producer statistics, so the next
        GroupNorm takes the exact two-pass slab kernel (small feature maps only)."""
        B, H, W, C = x.B, x.H, x.W, x.C
        Ho, Wo = (H // 2, W // 2) if mode == 1 else (H * 2, W * 2)
        r = self.buf(tag + '.rs', (B, Ho, Wo, C), torch.float32)
        (K.avgpool2_f32 if mode == 1 else K.upsample2_f32)(x.t, r, B, H, W, C)
        return Act(r, B, Ho, Wo, C)

    def _conv1_gn2_fused(self, tag, a1, B, Ho, Wo, Cin, Cout, conv1, norm2, emb, emb_off, emb_ld, scale_shift):
        """SiLU(GN2(conv1(a1) + bias [+ emb row]) [* (1 + scale) + shift]) in ONE launch -> the bf16 operand of conv2.
        The caller checked K.conv2d_gn_ok for this layer; a rejected launch raises (no per-layer fallback)."""
        w, b = self.w_conv(tag + '.c1', conv1)
        a2 = self.buf(tag + '.2.gn', (B, Ho, Wo, Cout), torch.bfloat16, temp=True)
        ws = {}
        if K.conv2d_gn_needs_workspace(Ho, Wo):     # several tiles per image: statistics / arrival counters, zeroed per forward
            ws = dict(xstats=self.stats_buf(tag + '.c1', B, Cout), xcount=self.stats_buf(tag + '.c1.cnt', B, 1))
        if scale_shift:
            K.conv2d_gn(a1, w, Cout, B, Ho, Wo, K.taps_3x3_s1(), a0_geom=(Cin, Ho, Wo, 1), gamma=norm2.weight,
                        beta=norm2.bias, groups=norm2.num_groups, eps=norm2.eps, out_norm=a2, bias=b,
                        scale=emb[:, emb_off:], shift=emb[:, emb_off + Cout:], ss_ld=emb_ld, **ws)
        else:
            K.conv2d_gn(a1, w, Cout, B, Ho, Wo, K.taps_3x3_s1(), a0_geom=(Cin, Ho, Wo, 1), gamma=norm2.weight,
                        beta=norm2.bias, groups=norm2.num_groups, eps=norm2.eps, out_norm=a2, bias=b,
                        rowadd=emb[:, emb_off:], rowadd_ld=emb_ld, **ws)
        return a2

    def next_gn_ok(self, B, H, W, C, next_gn) -> bool:
        """May the conv producing a [B, H, W, C] block output also apply `next_gn` = (GroupNorm, silu[, C_skip[, dead]]) of its
        consumer?  Every eligible layer takes the fused form: a per-layer policy derived from cold-cache ncu launch lists
        (no fusion for 32x32 layers that also stream a residual, no concat form below 16x16) was A/B-timed on one box
        and lost to `all layers` inside the sampling graph (DDIM-50 1195.0 vs 1197.0 images/s; none: 1178)."""
        if next_gn is None or not (self.fuse_gn1 and self.fuse_gn2) or self.split:
            return False
        if self.tape is not None and not self.fuse_gn1_train:
            return False
        norm, skip_C = next_gn[0], (next_gn[2] if len(next_gn) > 2 else 0)
        if skip_C == 0:
            return norm.num_channels == C and K.conv2d_gn_ok(B, H, W, C, norm.num_groups)
        # consumer normalises cat(this, skip): its groups must not straddle the boundary, and the skip's part must take
        # the per-thread-coefficient path of the streaming kernel in window mode (whole 8-channel columns)
        Ct = C + skip_C
        if not self.fuse_gn1_cat or norm.num_channels != Ct or Ct % norm.num_groups or Ct > 2048:
            return False
        cpg = Ct // norm.num_groups
        return (C % cpg == 0 and skip_C % 8 == 0 and C % 8 == 0 and (cpg in (1, 2, 4, 8) or cpg % 8 == 0)
                and K.conv2d_gn_ok(B, H, W, C, C // cpg))

    def _pre_buf(self, B, H, W, C, kind='pre', tag=None):
        """bf16 buffer for a producer-applied GroupNorm: read by the very next block only, so two per shape alternate.
        Training: the backward reads it (conv1's weight-gradient operand, the shortcut's), so every producer keeps its own."""
        if self.tape is not None:
            return self.buf(f'{tag}.{kind}', (B, H, W, C), torch.bfloat16)
        k = (kind, B, H, W, C)
        n = self._pp_count.get(k, 0)
        self._pp_count[k] = n + 1
        return self.buf(f'{kind}{n & 1}', (B, H, W, C), torch.bfloat16)

    def conv_block_out(self, tag, a, B, H, W, Cin, Cout, w, b, taps, next_gn, *, residual: Optional[Act] = None, sc_a=None,
                       sc_C=0, alg_macs=None) -> Act:
        """Last conv of a block with the consumer's GroupNorm fused: out = conv(a) + b (+ residual | + 1x1 shortcut
        K-blocks over sc_a) as fp32 NHWC with statistics, and out.pre = SiLU?(GN_next(out)) as bf16 (one launch)."""
        norm, silu = next_gn[0], next_gn[1]
        skip_C = next_gn[2] if len(next_gn) > 2 else 0
        # next_gn[3]: the consumer reads only `pre` / `praw` (output head, concatenating decoder block with a 1x1 shortcut)
        # and the output is no skip connection -> without a residual the fp32 form and its statistics are never written
        dead = len(next_gn) > 3 and next_gn[3] and residual is None and self.skip_dead_out and self.tape is None
        stats = self.stats_buf(tag, B, Cout)
        out = None if dead else self.buf(tag + '.out', (B, H, W, Cout), torch.float32)
        Ct = Cout + skip_C
        pre = self._pre_buf(B, H, W, Ct, tag=tag)
        rawc = self._pre_buf(B, H, W, Ct, 'praw', tag=tag) if skip_C else None
        groups = norm.num_groups if not skip_C else Cout // (Ct // norm.num_groups)
        ws = {}
        if K.conv2d_gn_needs_workspace(H, W):     # several tiles per image: the output statistics double as the exchange workspace
            ws = dict(xstats=stats, xcount=self.stats_buf(tag + '.cnt', B, 1))
        K.conv2d_gn(a, w, Cout, B, H, W, taps, a0_geom=(Cin, H, W, 1), gamma=norm.weight, beta=norm.bias,
                    groups=groups, eps=norm.eps, out_norm=pre, out_norm_ld=Ct if skip_C else 0, out_raw=rawc, bias=b,
                    silu=silu, out=out,
                    stats=None if (ws or dead) else stats, residual=None if residual is None else residual.t,
                    res_ld=0 if residual is None else residual.C, a1=sc_a,
                    a1_geom=(sc_C, H, W, 1) if sc_a is not None else None, alg_macs=alg_macs, **ws)
        res = Act(out, B, H, W, Cout, None if dead else stats)
        res.pre = (norm, bool(silu), pre, rawc) if skip_C else (norm, bool(silu), pre)
        return res

    def resblock_core(self, *args, **kwargs):
        with self.scope():       # the block's temporaries are released when it returns
            return self._resblock_core_body(*args, **kwargs)

    def _resblock_core_body(self, tag, x: Act, skip: Optional[Act], *, norm1, conv1, norm2, conv2, shortcut, emb, emb_off,
                      emb_ld, scale_shift: bool, resample: int = 0, dropout: Optional[nn.Dropout] = None,
                      emb_linear: Optional[nn.Linear] = None, next_gn=None) -> Act:
        """The ResBlock shared by all UNet families:
            h = conv1(resample(SiLU(GN1(cat(x, skip)))))            (+ emb row when not scale_shift)
            h = conv2(SiLU(GN2(h) [* (1 + scale) + shift]))          (dropout = identity in eval mode)
            out = h + shortcut(resample(cat(x, skip)))
        `emb` is the [rows, total] fp32 matrix of all per-block embedding projections, this block's columns start at
        emb_off (scale_shift: [scale | shift], 2*Cout columns).  shortcut: None (identity), a 1x1 conv (folded into
        conv2's GEMM as extra K-blocks) or a 3x3 conv (own launch, added as conv2's residual).
        resample: 0 none, 1 = 2x2 average pool, 2 = nearest 2x of both h and x (BigGAN-style up/down blocks).
        Reference: models/unet.py:30-43, models/unet_categorial_adagn.py:44-62, models/adm/unet.py:244-275,
        models/pesser/model.py:114-134."""
        B, H, W = x.B, x.H, x.W
        Cin = x.C + (skip.C if skip is not None else 0)
        Cout = conv1.out_channels
        sc1x1 = shortcut is not None and shortcut.kernel_size[0] == 1
        sc3x3 = shortcut is not None and not sc1x1
        if resample and (shortcut is not None or skip is not None):
            raise RuntimeError('up/down ResBlocks with a projection shortcut or a concatenated skip are not supported')
        a1, raw = self.gn(tag + '.1', x, skip, norm1, raw=sc1x1 or sc3x3, resample=resample)
        Ho, Wo = (H // 2, W // 2) if resample == 1 else (H * 2, W * 2) if resample == 2 else (H, W)
        res_x = x
        if resample:
            r = self.buf(tag + '.xr', (B, Ho, Wo, x.C), torch.float32, temp=True)
            (K.avgpool2_f32 if resample == 1 else K.upsample2_f32)(x.t, r, B, H, W, x.C)
            res_x = Act(r, B, Ho, Wo, x.C)
        drop_p, drop_seed = 0.0, 0
        if self.tape is not None and dropout is not None and dropout.p > 0 and self.model.training:
            drop_p, drop_seed = float(dropout.p), self.next_drop_seed()
        a2 = h = None
        if (self.fuse_gn2 and self.tape is None and drop_p == 0.0 and Cin % 64 == 0 and not self.split
                and K.conv2d_gn_ok(B, Ho, Wo, Cout, norm2.num_groups)):
            a2 = self._conv1_gn2_fused(tag, a1, B, Ho, Wo, Cin, Cout, conv1, norm2, emb, emb_off, emb_ld, scale_shift)
        elif scale_shift:
            h = self.conv3x3(tag + '.c1', a1, B, Ho, Wo, Cin, conv1, intermediate=True)
            a2, _ = self.gn(tag + '.2', h, None, norm2, scale=emb[:, emb_off:], shift=emb[:, emb_off + Cout:],
                            ss_ld=emb_ld, drop_p=drop_p, drop_seed=drop_seed)
        else:
            h = self.conv3x3(tag + '.c1', a1, B, Ho, Wo, Cin, conv1, rowadd=emb[:, emb_off:], rowadd_ld=emb_ld,
                             intermediate=True)
            a2, _ = self.gn(tag + '.2', h, None, norm2, drop_p=drop_p, drop_seed=drop_seed)
        fuse_next = not sc3x3 and Cout % 64 == 0 and self.next_gn_ok(B, Ho, Wo, Cout, next_gn)
        if fuse_next:       # conv2 (+ residual / 1x1 shortcut) also applies the consumer's GroupNorm (`next_gn`)
            w2, b2 = self.w_conv(tag + '.c2', conv2, shortcut if sc1x1 else None)
            if not sc1x1:
                assert skip is None and Cin == Cout
            out = self.conv_block_out(tag + '.c2', a2, B, Ho, Wo, Cout, Cout, w2, b2, K.taps_3x3_s1(), next_gn,
                                      residual=None if sc1x1 else res_x, sc_a=raw if sc1x1 else None, sc_C=Cin if sc1x1 else 0)
        elif sc1x1:
            out = self.conv3x3(tag + '.c2', a2, B, Ho, Wo, Cout, conv2, sc_a=raw, sc_C=Cin, sc_conv=shortcut)
        else:
            if sc3x3:
                res_x = self.conv3x3(tag + '.sc', raw, B, Ho, Wo, Cin, shortcut, temp=True)
            else:
                assert skip is None and Cin == Cout
            out = self.conv3x3(tag + '.c2', a2, B, Ho, Wo, Cout, conv2, residual=res_x)
        if self.tape is not None:
            self.tape.append(dict(kind='res', tag=tag, x=x, skip=skip, out=out, a1=a1, raw=raw, h=h, a2=a2, norm1=norm1,
                                  conv1=conv1, norm2=norm2, conv2=conv2, shortcut=shortcut, emb=emb, emb_off=emb_off,
                                  emb_ld=emb_ld, scale_shift=scale_shift, resample=resample, drop_p=drop_p,
                                  drop_seed=drop_seed, emb_linear=emb_linear))
        return out

    def resblock(self, tag, blk, x: Act, skip: Optional[Act], tproj, tproj_off, tproj_ld, next_gn=None) -> Act:
        """models/unet.py:30-43: conv(SiLU(GN(x))) + temb -> conv(SiLU(GN(h))) + shortcut(x).
        next_gn = (GroupNorm, silu) of the op that consumes this block's output as its ONLY GroupNorm input (the next
        ResBlock's norm1 without a concatenated skip, an attention block's norm, the output head's norm), or None."""
        return self.resblock_core(tag, x, skip, norm1=blk.blk1[0], conv1=blk.blk1[2], norm2=blk.blk2[0],
                                  conv2=blk.blk2[3],
                                  shortcut=blk.shortcut if isinstance(blk.shortcut, nn.Conv2d) else None,
                                  emb=tproj, emb_off=tproj_off, emb_ld=tproj_ld, scale_shift=False, dropout=blk.blk2[2],
                                  emb_linear=blk.proj[1], next_gn=next_gn)

    def resblock_adagn(self, tag, blk, x: Act, skip: Optional[Act], ss, ss_off, ss_ld, next_gn=None) -> Act:
        """models/unet_categorial_adagn.py:44-62 incl. the BigGAN-style up/down variants."""
        return self.resblock_core(tag, x, skip, norm1=blk.blk1[0], conv1=blk.blk1[2], norm2=blk.adagn.gn,
                                  conv2=blk.blk2[2],
                                  shortcut=blk.shortcut if isinstance(blk.shortcut, nn.Conv2d) else None,
                                  emb=ss, emb_off=ss_off, emb_ld=ss_ld, scale_shift=True,
                                  resample={'up': 2, 'down': 1}.get(blk.updown_kind, 0), dropout=blk.blk2[1],
                                  emb_linear=blk.adagn.proj[1], next_gn=next_gn)

    # ------------------------------------------------------------------------------------------
    # embedding path
    # ------------------------------------------------------------------------------------------
    def embed(self, T, y, B, pos_emb, lin1, lin2, class_embed, proj_linears):
        """Returns (proj [rows, sum_out] fp32, ld) where ld = 0 when one row serves the whole batch."""
        dev = self.device
        # one embedding row serves the whole batch when t is a stride-0 expand (sampling); training keeps per-sample rows
        uniform = (T.dim() == 1 and T.shape[0] == B and (B == 1 or T.stride(0) == 0)) and self.tape is None
        use_y = class_embed is not None and y is not None
        if uniform and not use_y:
            if self.emb_override is not None:   # the sampling runner filled this row for the current step
                return self.emb_override, 0
            self._embed_args = (pos_emb, lin1, lin2, proj_linears)
        rows = 1 if (uniform and not use_y) else B
        t_rows = (T[:1] if uniform and rows == 1 else T).to(torch.long).contiguous()
        if use_y:
            y = y.to(torch.long).contiguous()
        E = lin1.out_features
        freqs = self.const(('freqs', str(dev)), lambda: pos_emb.frequencies(dev).float().contiguous())
        m = self.m3
        emb = self.buf('emb', (rows, E), torch.float32)
        semb = self.buf('semb', (rows, m * E), torch.bfloat16)
        K.time_embed(t_rows, freqs, pos_emb.dim, E, bool(getattr(pos_emb, 'cos_first', False)), lin1.weight, lin1.bias,
                     lin2.weight, lin2.bias, emb,
                     y=y if use_y else None, class_embed=class_embed.weight if use_y else None,
                     out_silu_bf16=None if self.split else semb)
        if self.split:
            self._fp32_inference_only('embed')
            K.split_cast(emb, semb, rows, E, silu=True)       # exact SiLU, then [hi | lo | hi]
        w, b = self.w_tproj(proj_linears)
        total = w.shape[0]
        proj = self.buf('tproj', (rows, total), torch.float32)
        K.conv2d(semb, w, total, rows, 1, 1, K.taps_1x1(), a0_geom=(m * E, 1, 1, 1), bias=b, out=proj)
        if self.tape is not None:
            self.tape.append(dict(kind='embed', t=t_rows, y=y if use_y else None, rows=rows, E=E, total=total,
                                  pos_emb=pos_emb, freqs=freqs, lin1=lin1, lin2=lin2, class_embed=class_embed if use_y else None,
                                  linears=proj_linears, emb=emb, semb=semb, w=w))
            self.d_emb = self.buf('d_tproj', (rows, total), torch.float32)
            self.d_emb.zero_()
        return proj, (0 if rows == 1 and self.tape is None else total)

    def embed_rows(self, t_all):
        """Embedding projections of a whole timestep schedule at once: t_all [S] int64 -> [S, total] fp32, the rows that
        `embed` would produce one per call (same kernels, rows = S).  The sampling runner evaluates this once per
        `sample()` instead of a sinusoid + 2-layer MLP + projection GEMM in every one of the S steps."""
        pos_emb, lin1, lin2, proj_linears = self._embed_args
        dev = self.device
        S, E = t_all.numel(), lin1.out_features
        freqs = self.const(('freqs', str(dev)), lambda: pos_emb.frequencies(dev).float().contiguous())
        m = self.m3
        emb = torch.empty((S, E), dtype=torch.float32, device=dev)
        semb = torch.empty((S, m * E), dtype=torch.bfloat16, device=dev)
        K.time_embed(t_all.to(torch.long).contiguous(), freqs, pos_emb.dim, E, bool(getattr(pos_emb, 'cos_first', False)),
                     lin1.weight, lin1.bias, lin2.weight, lin2.bias, emb, out_silu_bf16=None if self.split else semb)
        if self.split:
            K.split_cast(emb, semb, S, E, silu=True)
        w, b = self.w_tproj(proj_linears)
        proj = torch.empty((S, w.shape[0]), dtype=torch.float32, device=dev)
        K.conv2d(semb, w, w.shape[0], S, 1, 1, K.taps_1x1(), a0_geom=(m * E, 1, 1, 1), bias=b, out=proj)
        return proj

    def first_conv(self, tag, conv: nn.Conv2d, X, next_gn=None) -> Act:
        """The Cin <= 4 input convolution (models/unet.py:72,123): NCHW fp32 image -> fp32 NHWC residual stream.
        Inference with Cout % 128 == 0: on the tensor cores -- the image becomes 64-channel bf16 pixels [hi | lo | hi | 0..]
        (K.first_split) against weights [w_hi | w_hi | w_lo | 0..] per tap, i.e. fp32-grade products, and the conv also
        applies `next_gn` (the first ResBlock's norm1) when eligible.  Otherwise (training, FP32 mode, narrow nets) the
        direct FP32 kernel."""
        B, Cin, H, W = X.shape
        Cout = conv.out_channels
        if (self.first_tc and self.tape is None and not self.split and Cout % 128 == 0 and 3 * Cin <= 64
                and conv.kernel_size == (3, 3) and X.dtype == torch.float32 and X.is_contiguous()):
            key = ('first_tc', tag)
            hit = self._pt.get(key)
            if hit is None:
                w = torch.zeros((Cout, 9 * 64), dtype=torch.bfloat16, device=self.device)
                w, = self.pack_table(key, (w,), [K.pack_entry_bytes(conv.weight, w, Cout, Cin, 9, 4, ld=9 * 64, tap_ld=64)])
            else:
                w, = hit[0]
            a = self.buf(tag + '.split', (B, H, W, 64), torch.bfloat16)
            K.first_split(X, a)
            if self.next_gn_ok(B, H, W, Cout, next_gn):
                return self.conv_block_out(tag, a, B, H, W, 64, Cout, w, conv.bias, K.taps_3x3_s1(), next_gn,
                                           alg_macs=float(B) * H * W * Cout * 9 * Cin)
            h0 = self.buf(tag + '.out', (B, H, W, Cout), torch.float32)
            st0 = self.stats_buf(tag, B, Cout)
            K.conv2d(a, w, Cout, B, H, W, K.taps_3x3_s1(), a0_geom=(64, H, W, 1), bias=conv.bias,
                     alg_macs=float(B) * H * W * Cout * 9 * Cin, out=h0, out_mode=K.OUT_F32_NHWC, stats=st0)
            return Act(h0, B, H, W, Cout, st0)
        h0 = self.buf(tag + '.out', (B, H, W, conv.out_channels), torch.float32)
        st0 = self.stats_buf(tag, B, conv.out_channels)
        K.conv3x3_first(X, conv.weight, conv.bias, h0, st0)
        res = Act(h0, B, H, W, conv.out_channels, st0)
        if self.tape is not None:
            self.tape.append(dict(kind='first', tag=tag, X=X, out=res, conv=conv))
        return res

    def head(self, *args, **kwargs):
        with self.scope():       # the block's temporaries are released when it returns
            return self._head_body(*args, **kwargs)

    def _head_body(self, tag, h: Act, norm: nn.GroupNorm, conv: nn.Conv2d, out):
        """GroupNorm -> SiLU -> 3x3 conv to the few output channels, written as the reference's NCHW fp32."""
        a, _ = self.gn(tag, h, None, norm)
        if out is None:
            out = torch.empty((h.B, conv.out_channels, h.H, h.W), dtype=torch.float32, device=a.device)
        self.conv3x3(tag + '.c', a, h.B, h.H, h.W, h.C, conv, out_mode=K.OUT_F32_NCHW, out=out)
        if self.tape is not None:
            self.tape.append(dict(kind='head', tag=tag, x=h, a=a, norm=norm, conv=conv))
        return out

    def check_input(self, X, T, in_channels):
        """Validates the network input and counts this forward against the arena of its shape: the training tape
        (models/backward.py) refers to arena buffers instead of saving copies, so its backward checks that no other
        forward of the same shape ran in between."""
        self._cur_key = tuple(X.shape) + (self.lane,)
        self.pingpong = False
        self._pp_count.clear()
        st = self._scratch.get(self._cur_key)
        if st is not None:
            st['pos'][0], st['pos'][1] = 0, 0
        self._fwd_gen[self._cur_key] = self._fwd_gen.get(self._cur_key, 0) + 1
        for pool, used in self._stats_pools.get(self._cur_key, ()):
            pool[:used].zero_()
        if X.dim() != 4 or X.shape[1] != in_channels:
            raise RuntimeError(f'expected input [B, {in_channels}, H, W], got {tuple(X.shape)}')
        if not X.is_cuda:
            raise RuntimeError('b200diff models run on CUDA tensors only (no CPU fallback)')
        if X.dtype != torch.float32:
            raise RuntimeError(f'expected float32 input, got {X.dtype}')
        return X.contiguous()
