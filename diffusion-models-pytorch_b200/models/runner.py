"""CUDA-graph replay of the sampling loop's per-timestep work.

The reference loop (diffusions/ddpm.py:263-281, ddim.py:161-191) launches ~480 eager ops per timestep.  Here one
timestep = {UNet forward (two for classifier-free guidance), the per-step noise draw, the fused sampler update}
is captured ONCE per (batch shape, conditioning layout) into a CUDA graph whose scalar inputs -- the timestep
and the row of sampler coefficients -- live in device memory; each loop iteration then costs two 8/48-byte
device-to-device copies plus one graph launch, with no host<->device synchronisation anywhere in the loop.
The noise is drawn with torch's Philox generator inside the graph, once per step like the reference
(ddim.py:76 / ddpm.py:251), so seeded runs consume the RNG stream identically.
"""
import weakref

import torch
import tqdm

import b200diff as K


class SamplingRunner:
    def __init__(self, model, diffuser):
        self.model = model
        self._diffuser = weakref.ref(diffuser)   # the model's runner table is keyed weakly by the diffuser
        self._graphs = {}

    @property
    def diffuser(self):
        return self._diffuser()

    # ------------------------------------------------------------------------------------------
    def _eager(self, init_noise, tqdm_kwargs, model_kwargs, guidance_scale, uncond_conditioning):
        d = self.diffuser
        sample = None
        if guidance_scale is None:
            loop = d.sample_loop(self.model, init_noise, tqdm_kwargs, model_kwargs)
        else:
            loop = d.sample_loop(self.model, init_noise, uncond_conditioning, tqdm_kwargs, model_kwargs)
        for out in loop:
            sample = out['sample']
        return sample

    def run(self, init_noise, tqdm_kwargs, model_kwargs, guidance_scale, uncond_conditioning):
        d = self.diffuser
        cfg = guidance_scale is not None
        tqdm_kwargs = dict() if tqdm_kwargs is None else tqdm_kwargs
        if cfg:
            if d.cond_kwarg not in model_kwargs.keys():
                raise ValueError(f'Condition argument `{d.cond_kwarg}` not found in model_kwargs.')
        else:
            model_kwargs = dict() if model_kwargs is None else model_kwargs
        # graph mode supports tensor / None keyword arguments only
        graphable = (init_noise.is_cuda and not torch.cuda.is_current_stream_capturing()
                     and all(v is None or torch.is_tensor(v) for v in model_kwargs.values())
                     and (uncond_conditioning is None or torch.is_tensor(uncond_conditioning)))
        if not graphable:
            return self._eager(init_noise, tqdm_kwargs, model_kwargs, guidance_scale, uncond_conditioning)

        eng = self.model.engine
        eng.refresh()
        layout = tuple(sorted((k, None if v is None else (tuple(v.shape), v.dtype)) for k, v in model_kwargs.items()))
        key = (tuple(init_noise.shape), init_noise.device, cfg, layout,
               None if uncond_conditioning is None else tuple(uncond_conditioning.shape),
               d.objective, d.clip_denoised, d.var_type, getattr(d, 'eta', None), guidance_scale)
        g = self._graphs.get(key)
        if g is None or g['sig'] != eng._sig:
            g = self._capture(init_noise, model_kwargs, cfg, guidance_scale, uncond_conditioning)
            g['sig'] = eng._sig
            self._graphs[key] = g

        pairs = d._step_pairs()
        t_table = torch.tensor([t for t, _ in pairs], dtype=torch.long, device=init_noise.device)
        coef_table = d._coef_table(pairs)
        g['x'].copy_(init_noise)
        if g.get('x_dup') is not None:      # batched guidance: the unconditional half of the 2B batch sees the same x_t
            g['x_dup'].copy_(init_noise)
        for k, v in model_kwargs.items():
            if v is not None:
                g['kw'][k].copy_(v)
        if cfg and uncond_conditioning is not None:
            g['uncond'].copy_(uncond_conditioning)
        emb_table = None
        if g['emb'] is not None:   # embedding projections of the whole schedule in one batched evaluation
            with torch.no_grad():
                emb_table = eng.embed_rows(t_table)
        pbar = tqdm.tqdm(total=len(pairs), **tqdm_kwargs)
        for i in range(len(pairs)):
            g['t'].copy_(t_table[i:i + 1])
            if emb_table is not None:
                g['emb'].copy_(emb_table[i:i + 1])
            g['coef'].copy_(coef_table[i])
            g['graph'].replay()
            K.GRAPH_LAUNCHES += g['kernels']
            pbar.update(1)
        pbar.close()
        return g['x'].clone()

    # ------------------------------------------------------------------------------------------
    def _capture(self, init_noise, model_kwargs, cfg, guidance_scale, uncond_conditioning):
        d, model = self.diffuser, self.model
        dev = init_noise.device
        B = init_noise.shape[0]
        rng = torch.cuda.get_rng_state(dev)   # building the graph must not consume the caller's RNG stream
        st = {
            'x': torch.zeros_like(init_noise),
            't': torch.zeros(1, dtype=torch.long, device=dev),
            'coef': d._coef_row(*d._step_pairs()[0]).clone(),
            'kw': {k: (None if v is None else v.clone()) for k, v in model_kwargs.items()},
            'uncond': None if uncond_conditioning is None else uncond_conditioning.clone(),
        }
        out_ch = getattr(model, 'out_channels', init_noise.shape[1])
        learned = d.var_type == 'learned_range' and d._uses_learned_var()
        # Classifier-free guidance as ONE forward over [x_t ; x_t] with labels [y ; -1] (a negative label adds no class
        # embedding = the reference's y=None branch, models/unet_categorial_adagn.py:172-174) instead of two B-sized
        # forwards (ddim.py:177-183): half the launches per step and twice the tiles per launch at the 8x8 / 4x4 levels.
        # Only for models that declare `cfg_batch_ok` (one network serves both branches), the reference's standard
        # call pattern (class labels in model_kwargs[cond_kwarg], uncond_conditioning=None); B200_CFG_BATCH=0 disables.
        y_cond = st['kw'].get(d.cond_kwarg) if cfg else None
        batched = (cfg and getattr(model, 'cfg_batch_ok', False) and uncond_conditioning is None and len(st['kw']) == 1
                   and torch.is_tensor(y_cond) and y_cond.dim() == 1 and y_cond.shape[0] == B
                   and y_cond.dtype in (torch.int64, torch.int32)
                   and __import__('os').environ.get('B200_CFG_BATCH', '1') != '0')
        if batched:
            x2 = torch.zeros((2 * B,) + tuple(init_noise.shape[1:]), dtype=init_noise.dtype, device=dev)
            y2 = torch.full((2 * B,), -1, dtype=torch.long, device=dev)
            out2 = torch.empty((2 * B, out_ch) + tuple(init_noise.shape[2:]), dtype=torch.float32, device=dev)
            st['x'], st['x_dup'] = x2[:B], x2[B:]
            st['kw'] = {d.cond_kwarg: y2[:B]}        # run() copies the caller's labels into the conditional half
            st['out_c'], st['out_u'] = out2[:B], out2[B:]
        else:
            st['out_c'] = torch.empty((B, out_ch) + tuple(init_noise.shape[2:]), dtype=torch.float32, device=dev)
            st['out_u'] = torch.empty_like(st['out_c']) if cfg else None
        st['cfg_batched'] = batched

        # Two half-batch forwards as parallel branches of the step graph (B200_LANES=2, experiment knob): the branches use
        # disjoint buffer sets (Engine.lane) and share the packed weights; one branch's HBM-bound GroupNorm / epilogue
        # phases can then overlap the other's tensor-bound mainloops.  Unconditional models without keyword tensors only.
        lanes = 2 if (__import__('os').environ.get('B200_LANES', '1') == '2' and not cfg and B % 2 == 0 and B >= 32
                      and not st['kw'] and hasattr(model, 'engine') and hasattr(model.engine, 'lane')) else 1
        st['lanes'] = lanes
        side2 = torch.cuda.Stream(device=dev) if lanes == 2 else None

        def forward_lanes():
            half = B // 2
            cur = torch.cuda.current_stream(dev)
            fork = torch.cuda.Event()
            fork.record(cur)
            side2.wait_event(fork)
            eng_ = model.engine
            try:
                eng_.lane = 0
                model(st['x'][:half], st['t'].expand(half), out=st['out_c'][:half])
                with torch.cuda.stream(side2):
                    eng_.lane = 1
                    model(st['x'][half:], st['t'].expand(half), out=st['out_c'][half:])
            finally:
                eng_.lane = 0
            join = torch.cuda.Event()
            join.record(side2)
            cur.wait_event(join)

        def step():
            if batched:
                model(x2, st['t'].expand(2 * B), out=out2, **{d.cond_kwarg: y2})
            elif lanes == 2:
                forward_lanes()
            else:
                tb = st['t'].expand(B)
                model(st['x'], tb, out=st['out_c'], **st['kw'])
                if cfg:
                    ukw = dict(st['kw'])
                    ukw[d.cond_kwarg] = st['uncond']
                    model(st['x'], tb, out=st['out_u'], **ukw)
            noise = torch.randn_like(st['x'])
            K.sampler_step(st['out_c'], st['x'], st['coef'], objective=d.objective, clip=d.clip_denoised,
                           learned_range=learned, noise=noise, model_out_uncond=st['out_u'],
                           guidance_scale=guidance_scale if cfg else 1.0, sample=st['x'])
            if batched:
                st['x_dup'].copy_(st['x'])

        # warm-up on a side stream (fills the arena, packs weights, sets kernel attributes); RNG state preserved
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side), torch.no_grad():
            step()
        torch.cuda.current_stream(dev).wait_stream(side)
        # models whose embedding depends on the timestep only (no class label): the captured forward reads the
        # projections of the current step from st['emb'], filled per step from a table computed once per sample()
        eng = model.engine
        st['emb'] = None
        if (not cfg and getattr(eng, '_embed_args', None) is not None and hasattr(eng, 'embed_rows')
                and all(v is None for v in st['kw'].values())):
            total = sum(l.out_features for l in eng._embed_args[3])
            st['emb'] = torch.zeros((1, total), dtype=torch.float32, device=dev)
        graph = torch.cuda.CUDAGraph()
        n0 = K.direct_launch_count()
        eng.emb_override = st['emb']
        try:
            with torch.no_grad(), torch.cuda.graph(graph):
                step()
        finally:
            eng.emb_override = None
        st['kernels'] = K.direct_launch_count() - n0   # libb200diff kernels captured per timestep
        torch.cuda.set_rng_state(rng, dev)
        st['graph'] = graph
        return st
