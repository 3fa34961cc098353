"""Native ConvLSTM classifier runner (pt/models/convolution_lstm.py:63-132 + pt/models/CLSTM_4.py:69-85)
as a fixed sequence of libivf launches, forward and BPTT data-gradient, CUDA-graph capturable.

Schedule (results identical to the reference's step-major loop, :99-129, because layer l at step t
depends on layer l-1 at step t only):
  per layer:  x-convolution of ALL T steps as one implicit GEMM (4 gates side by side, bias in the
              epilogue)  ->  T sequential steps { h-convolution accumulated into the gate
              pre-activations by the conv epilogue (skipped at t=0: zero state); fused gate kernel }
              ->  eval-BatchNorm + MaxPool2d(2) of all steps in one launch.
  classifier: Linear(+softmax) on the last effective step (use_entire_seq=False).
Frames are time-major ([t][b]...) so that a step's batch is contiguous.

bf16 mode: the stride-2 5x5 x-convolutions run space-to-depth (stride-1 3x3 over 4*cin channels, the
perturbation kernel / the pool kernel write that layout directly); hidden sizes below 8 are padded
to 8 with zero weights (padded channels stay exactly 0: gates 0.5/0.5/0/0.5 on a zero state).
"""
import torch

from . import _lib, ops, tune
from ._lib import PFMT_S2D2_BF16, PFMT_TBHWC_F32
from .engine import _dev32, pack, strip_module_prefix
from .ops import Act

GATES = ("i", "f", "c", "o")


class CLSTMEngine:
    def __init__(self, state_dict, batch, clip, hidden, layers, num_classes, kernel=5, conv_stride=2, mode="bf16",
                 softmax=False, batch_norm=True, effective_step=(7, 15, 23, 31), device=None, in_channels=3):
        assert mode in ("bf16", "fp32")
        if kernel != 5 or conv_stride != 2:
            raise _lib.IvfError("native ConvLSTM supports kernel 5 / conv_stride 2 (the reference configs)")
        self.mode = mode
        self.dtype = torch.bfloat16 if mode == "bf16" else torch.float32
        self.device = dev = torch.device(device if device is not None else "cuda")
        _lib.handle(dev)
        sd = {k: v.detach().to(device=dev, dtype=torch.float32) for k, v in strip_module_prefix(state_dict).items()
              if v.is_floating_point()}
        self.B, (self.T, self.H, self.W) = batch, clip
        self.C, self.L, self.hid, self.softmax = in_channels, layers, hidden, bool(softmax)
        # the reference collects outputs only for steps that occur (convolution_lstm.py:126 `if step in
        # self.effective_step` inside `for step in range(self.step)`), so output[-1] is the last effective step
        # below T
        self.eff = [int(e) for e in effective_step if 0 <= int(e) < clip[0]]
        if not self.eff:
            raise _lib.IvfError("ConvLSTM: no effective_step below step=%d (%s): the reference's output list "
                                "would be empty" % (clip[0], list(effective_step)))
        self.generation = 0
        bf = mode == "bf16"
        he = max(8, (hidden + 7) // 8 * 8) if bf else hidden  # padded hidden size
        # IVF_CLSTM_FUSED=1 (bf16): the recurrent step as ONE kernel (ivf_conv3d_lstm: h-convolution with the gates, the
        # c/h update and the gate activations in its epilogue, unit-major channels).  Built, parity-tested and
        # measured on a B200 (C3, 8 clips): 206 launches and 3.77 ms per step against 268 launches and 3.70 ms for the
        # convolution + gate-kernel pair - the gate math (4 expf + 2 tanhf per hidden unit) lengthens the epilogue of
        # a one-to-two-wave convolution by more than the separate bandwidth-bound gate kernel costs - so it is opt-in.
        import os
        # IVF_CLSTM_WAVE (bf16, default on): the layers run as a wavefront on one stream each - layer l's step t starts
        # as soon as layer l-1 has produced its step t (forward) / layer l+1 its step t (backward) - instead of one
        # whole recurrence after the other, so that the upper layers' small, latency-bound launches (15x20 maps) hide
        # behind the first layer's.  A convolution CTA owns its SM (200 KB of shared memory), so this only works when
        # an upper layer's launch fits on the SMs the first layer leaves (120 CTAs of 148 at 8 clips): the upper
        # layers get the plan with the fewest CTAs (_small_plan: 24).  Measured on a B200 (C3): 3.45 -> 3.08 ms per
        # step at 8 clips, 5.39 -> 4.82 ms at 16, 8.71 -> 8.51 ms at 32; without the small plans nothing was gained
        # (the second layer's 64-CTA launches waited for the first layer's to drain).
        self.wave = bf and layers > 1 and os.environ.get("IVF_CLSTM_WAVE", "1") != "0"
        self._lanes, self._ev_f, self._ev_b = None, None, None
        self.fused_mode = os.environ.get("IVF_CLSTM_FUSED", "0") if bf else "0"  # "2": small maps only (per layer)
        self.unit_major = self.fused_mode == "1"
        self.he = he
        B, T = batch, clip[0]
        N = T * B
        if self.H % 2 or self.W % 2:
            raise _lib.IvfError("native ConvLSTM needs even frame sizes")

        self.bn_scale = ops.zeros((he,), torch.float32, dev)  # padded channels: scale 0, shift 0
        self.bn_shift = ops.zeros((he,), torch.float32, dev)
        lib = _lib.load()
        if batch_norm:  # BatchNorm2d(eps=1e-05), convolution_lstm.py:85
            _lib.check(lib.ivf_bn_fold(_lib.handle(dev), *(_lib.ptr(sd["clstm.bn." + k]) for k in
                                                            ("weight", "bias", "running_mean", "running_var")),
                                       1e-5, None, hidden, _lib.ptr(self.bn_scale), _lib.ptr(self.bn_shift),
                                       _lib.stream_ptr(dev)), "ivf_bn_fold")
        else:
            _lib.check(lib.ivf_bn_fold(_lib.handle(dev), None, None, None, None, 0.0, None, hidden,
                                       _lib.ptr(self.bn_scale), _lib.ptr(self.bn_shift), _lib.stream_ptr(dev)),
                       "ivf_bn_fold")

        # ---- input operand of layer 0
        if bf:
            self.in_fmt = PFMT_S2D2_BF16
            self.xin = Act.empty(N, 1, self.H // 2, self.W // 2, 16, torch.bfloat16, dev, zero=True)
            self.g_xin = Act.empty(N, 1, self.H // 2, self.W // 2, 16, torch.bfloat16, dev, zero=True)
        else:
            self.in_fmt = PFMT_TBHWC_F32
            self.xin = Act.empty(N, 1, self.H, self.W, in_channels, torch.float32, dev)
            self.g_xin = Act.empty(N, 1, self.H, self.W, in_channels, torch.float32, dev)

        self.layers = []
        hin, win, cin = self.H, self.W, in_channels
        x_act, gx_in = self.xin, self.g_xin
        for l in range(layers):
            p = "clstm.cell%d." % l
            cin_real = cin if l == 0 else hidden
            cin_eff = cin_real if l == 0 else he
            # the four gates side by side (i, f, c, o), each padded from `hidden` to `he` channels
            wxs = [sd[p + "Wx%s.weight" % g].contiguous() for g in GATES]
            whs = [sd[p + "Wh%s.weight" % g].contiguous() for g in GATES]
            # gate-major [i.. | f.. | c.. | o..], or unit-major [i0 f0 c0 o0 i1 ...] when the gates run in the recurrent
            # convolution's epilogue (a 16-column accumulator chunk then holds four whole hidden units)
            # "2": fuse the layers whose recurrent convolution is less than one wave of tiles (a launch of pure
            # latency either way: the fused form saves the gate kernel's launch)
            um = self.unit_major or (self.fused_mode == "2" and B * (hin // 2) * (win // 2) <= 128 * 148)
            gate_offs = list(range(4)) if um else [gi * he for gi in range(4)]
            bias_host = torch.zeros(4 * he)
            for gi, g in enumerate(GATES):
                b_g = sd[p + "Wx%s.bias" % g].detach().float().cpu()
                if um:
                    bias_host[gi:4 * hidden:4] = b_g
                else:
                    bias_host[gi * he:gi * he + hidden] = b_g
            bias = bias_host.to(dev)
            ho, wo = hin // 2, win // 2
            if (hin + 4 - 5) // 2 + 1 != ho or (win + 4 - 5) // 2 + 1 != wo:
                raise _lib.IvfError("ConvLSTM layer %d: odd input %dx%d (the reference asserts here too)" % (l, hin, win))
            last = l == layers - 1
            s2d_out = bf and not last
            if s2d_out and ((ho // 2) % 2 or (wo // 2) % 2):
                raise _lib.IvfError("bf16 ConvLSTM needs an even pooled map between layers; use mode='fp32'")
            ones = torch.empty(4 * he, dtype=torch.float32, device=dev)
            _lib.check(lib.ivf_fill_u32(_lib.handle(dev), _lib.ptr(ones), 16 * he, 0x3F800000, _lib.stream_ptr(dev)),
                       "ivf_fill_u32")  # 1.0f
            rec = dict(l=l, ho=ho, wo=wo, hin=hin, win=win, bias=bias, ones=ones, um=um)
            gk = dict(co_offs=gate_offs, co_total=4 * he, co_stride=4 if um else 1)
            if bf:  # stride-2 5x5 as a stride-1 3x3 over the 2-D space-to-depth record (12 -> 16 channels at l = 0)
                xk = dict(s2d=(1, 2, 2), ci_stride=cin_eff, ceff_total=16 if l == 0 else None, **gk)
                rec.update(wx_f=pack(wxs, mode, **xk), wx_d=pack(wxs, mode, dgrad=True, **xk), xk=(1, 3, 3),
                           xs=(1, 1, 1), xpf=(0, 1, 1), xdpf=(0, 1, 1))
            else:
                rec.update(wx_f=pack(wxs, mode, ci_stride=cin_eff, **gk),
                           wx_d=pack(wxs, mode, dgrad=True, ci_stride=cin_eff, **gk), xk=(1, 5, 5), xs=(1, 2, 2),
                           xpf=(0, 2, 2))
            rec.update(wh_f=pack(whs, mode, ci_stride=he, **gk), wh_d=pack(whs, mode, dgrad=True, ci_stride=he, **gk))
            rec["_src"] = (wxs, whs)
            rec["x"], rec["g_x"] = x_act, gx_in
            rec["gx"] = Act.empty(N, 1, ho, wo, 4 * he, torch.float32, dev)            # gate pre-activations
            rec["h"] = Act.empty(N, 1, ho, wo, he, self.dtype, dev, zero=True)
            rec["c"] = ops.zeros((N, ho, wo, he), torch.float32, dev)
            rec["gact"] = ops.zeros((N, ho * wo, 4 * he), torch.float32, dev)
            rec["argmax"] = ops.zeros((N, ho // 2, wo // 2, he), torch.uint8, dev)
            if s2d_out:
                rec["pooled"] = Act.empty(N, 1, ho // 4, wo // 4, 4 * he, self.dtype, dev, zero=True)
            else:
                rec["pooled"] = Act.empty(N, 1, ho // 2, wo // 2, he, self.dtype, dev, zero=True)
            rec["s2d_out"] = s2d_out
            rec["g_pooled"] = rec["pooled"].like(zero=True)
            rec["dH"] = ops.zeros((N, ho, wo, he), torch.float32, dev)
            rec["dc"] = ops.zeros((B, ho, wo, he), torch.float32, dev)
            rec["dpre"] = Act.empty(N, 1, ho, wo, 4 * he, self.dtype, dev, zero=True)
            self.layers.append(rec)
            x_act, gx_in = rec["pooled"], rec["g_pooled"]
            hin, win = ho // 2, wo // 2
        self.fh, self.fw = hin, win

        # ---- classifier: Linear over the NCHW-flattened last effective step (CLSTM_4.py:78-80), columns
        # permuted once to our channels-last flattening
        full_sd = strip_module_prefix(state_dict)
        if "endFC.weight" in full_sd:
            wfc = full_sd["endFC.weight"].detach().float().cpu()  # permuted on the host
            ncls = wfc.shape[0]
            if wfc.shape[1] != hidden * hin * win:
                raise _lib.IvfError("endFC expects %d features, the stack produces %d (use_entire_seq is not supported)"
                                    % (wfc.shape[1], hidden * hin * win))
            wp = torch.zeros((ncls, hin * win, he))
            wp[:, :, :hidden] = wfc.view(ncls, hidden, hin * win).permute(0, 2, 1)
            self.w_fc = wp.reshape(ncls, -1).contiguous().to(dev)
            self.b_fc = sd["endFC.bias"].contiguous()
        else:  # the recurrent stack on its own (models.convolution_lstm.ConvLSTM.forward): no classifier
            ncls, self.w_fc, self.b_fc = 1, None, None
        self.num_classes = ncls
        self.logits = ops.zeros((B, ncls), torch.float32, dev)
        self.probs = ops.zeros((B, ncls), torch.float32, dev)
        self.dprobs = ops.zeros((B, ncls), torch.float32, dev)
        self.x = ops.zeros((B, in_channels, T, self.H, self.W), torch.float32, dev)
        self.dm = ops.zeros((B, T), torch.float32, dev)
        self.zero_mask = ops.zeros((B, T), torch.float32, dev)
        self.g_feat_raw = ops.zeros((B, hin * win * he), torch.float32, dev)
        self.head_ws = ops.head_workspace(B, hin * win * he, ncls, dev)

        # measured tile plans for the recurrent convolutions: each is a 15-35 us launch repeated T-1 times per
        # layer and direction with its operands resident in L2 (what the isolated measurement sees)
        if mode == "bf16" and tune.enabled() and T > 1:
            from .engine import ConvOp
            with torch.cuda.device(dev):
                for rec in self.layers:
                    dH = Act(rec["dH"], T * B, 1, rec["ho"], rec["wo"], he, 0, he)
                    gx1, dprev = self._step(rec["gx"], 1), self._step(dH, 0)
                    fwd = ConvOp(self._step(rec["h"], 0), rec["wh_f"], gx1, (1, 5, 5), (1, 1, 1), (0, 2, 2), acc_in=gx1)
                    bwd = ConvOp(self._step(rec["dpre"], 1), rec["wh_d"], dprev, (1, 5, 5), (1, 1, 1), (0, 2, 2),
                                 acc_in=dprev)
                    fwd.tune(dev, min_ms=0.012)
                    bwd.tune(dev, min_ms=0.012)
                    rec["plan_hf"], rec["plan_hd"] = fwd.plan, bwd.plan
                    ops.fill_zero(rec["gx"].buf)
                    ops.fill_zero(rec["dH"])
                torch.cuda.synchronize(dev)

    # ------------------------------------------------------------------ helpers
    def _step(self, act, t):
        """Act view of step t's B frames of a time-major [T*B,...] Act."""
        B = self.B
        per = act.d * act.h * act.w * act.ld
        buf = act.buf.view(-1)[t * B * per:(t + 1) * B * per]
        return Act(buf, B, act.d, act.h, act.w, act.ld, act.coff, act.c)

    def _feat(self, act, t):
        a = self._step(act, t)
        k = a.h * a.w * a.ld
        return Act(a.buf, self.B, 1, 1, 1, k, 0, k)

    def set_input(self, x):
        """x: [B,3,T,H,W] in the loader's layout, host or device, fp32 (0..255) or uint8 (frames as decoded: they
        cross PCIe as bytes and are converted on the device); copied into the static buffer."""
        assert tuple(x.shape) == (self.B, self.C, self.T, self.H, self.W), (x.shape,)
        if x.dtype == torch.uint8:
            if getattr(self, "x_u8", None) is None:
                self.x_u8 = torch.empty(self.x.shape, dtype=torch.uint8, device=self.device)
            self.x_u8.copy_(x, non_blocking=True)
            ops.u8_to_f32(self.x_u8, self.x)
        else:
            self.x.copy_(x, non_blocking=True)

    @_lib.on_device
    def set_targets(self, targets):
        self.generation += 1
        self._targets = ops.as_int32_targets(targets, self.device)
        ops.one_hot(self._targets, self.dprobs)

    # ------------------------------------------------------------------ forward
    # ------------------------------------------------------------------ wavefront schedule (bf16)
    # upper layers in the wavefront: single CTAs, one accumulator per tile, all of N in one tile - the FEWEST CTAs a
    # launch can have (24 on the 15x20 map at 8 clips), so that it fits on the SMs the first layer's launch leaves
    _small_plan = (1, 1, 2, 1, 1, 1)

    def _wave_setup(self):
        if self._lanes is None:
            for rec in self.layers[1:]:
                rec["plan_hf"] = rec["plan_hd"] = self._small_plan
            # first layer's recurrent data gradient: where the plan table has no entry the cost model picks CTA pairs
            # (128 CTAs at 8 clips), which leave the upper layer's launches no room; single CTAs with two
            # double-buffered accumulators measured 3.07 -> 2.94 ms per step at 8 clips
            import os
            env = os.environ.get("IVF_CLSTM_L0_DGRAD_PLAN")  # diagnostics: "kwm,mt,acc,ncta,ntiles,ds" or "auto"
            if env and env != "auto":
                self.layers[0]["plan_hd"] = tuple(int(v) for v in env.split(","))
            elif not env and self.layers[0].get("plan_hd") is None:
                self.layers[0]["plan_hd"] = (1, 2, 2, 1, 1, 1)
        if self._lanes is None:
            L, T = len(self.layers), self.T
            self._lanes = [None] + [torch.cuda.Stream(self.device) for _ in range(L - 1)]
            self._ev_f = [[torch.cuda.Event() for _ in range(T)] for _ in range(L)]
            self._ev_b = [[torch.cuda.Event() for _ in range(T)] for _ in range(L)]

    def _cell_step(self, rec, t):
        """h-convolution (accumulated onto the x-convolution's pre-activations) + gates of layer `rec`, step t."""
        B, he = self.B, self.he
        m = B * rec["ho"] * rec["wo"]
        gx_t = self._step(rec["gx"], t)
        c_prev = rec["c"][(t - 1) * B:t * B] if t > 0 else None
        c_next, gact_t = rec["c"][t * B:(t + 1) * B], rec["gact"][t * B:(t + 1) * B].view(m, 4 * he)
        if t > 0 and rec["um"]:
            ops.conv_lstm_step(self._step(rec["h"], t - 1), rec["wh_f"], gx_t, c_prev, c_next, self._step(rec["h"], t),
                               gact_t, (1, 5, 5), (0, 2, 2), plan=rec.get("plan_hf"))
            return
        if t > 0:
            ops.conv3d(self._step(rec["h"], t - 1), rec["wh_f"], gx_t, (1, 5, 5), (1, 1, 1), (0, 2, 2), acc_in=gx_t,
                       plan=rec.get("plan_hf"))
        ops.clstm_gates_fwd(gx_t.buf.view(m, 4 * he), c_prev, c_next, self._step(rec["h"], t).buf, gact_t,
                            unit_major=rec["um"])

    def _pool_step_fwd(self, rec, t):
        B, he = self.B, self.he
        h_t = rec["h"].buf.view(self.T * B, rec["ho"], rec["wo"], he)[t * B:(t + 1) * B]
        ops.bn_pool2d_fwd(h_t, self.bn_scale, self.bn_shift, self._step(rec["pooled"], t).buf,
                          rec["argmax"][t * B:(t + 1) * B], s2d=rec["s2d_out"])

    def _forward_wave(self):
        self._wave_setup()
        L, T = len(self.layers), self.T
        main = torch.cuda.current_stream(self.device)
        r0 = self.layers[0]
        ops.conv3d(r0["x"], r0["wx_f"], r0["gx"], r0["xk"], r0["xs"], r0["xpf"], scale=r0["ones"], shift=r0["bias"])
        # issue order = wavefront order, so that a single hardware queue would still see runnable work first
        for wave in range(T + L - 1):
            for l in range(min(L - 1, wave), -1, -1):
                t = wave - l
                if t < 0 or t >= T:
                    continue
                rec = self.layers[l]
                st = main if l == 0 else self._lanes[l]
                with torch.cuda.stream(st):
                    if l > 0:
                        st.wait_event(self._ev_f[l - 1][t])
                        ops.conv3d(self._step(rec["x"], t), rec["wx_f"], self._step(rec["gx"], t), rec["xk"], rec["xs"],
                                   rec["xpf"], scale=rec["ones"], shift=rec["bias"], plan=self._small_plan)
                    self._cell_step(rec, t)
                    self._pool_step_fwd(rec, t)
                    if l < L - 1:
                        self._ev_f[l][t].record(st)
        for l in range(1, L):
            main.wait_stream(self._lanes[l])

    def _backward_wave(self):
        """BPTT as a wavefront: the top layer leads, layer l's step t follows layer l+1's step t."""
        L, T, B, he = len(self.layers), self.T, self.B, self.he
        main = torch.cuda.current_stream(self.device)
        for l in range(1, L):
            self._lanes[l].wait_stream(main)  # head' (and everything before it) is done

        def stream_of(l):
            return main if l == 0 else self._lanes[l]

        top = self.layers[-1]
        with torch.cuda.stream(stream_of(L - 1)):
            ops.bn_pool2d_bwd(top["g_pooled"].buf, top["argmax"], self.bn_scale, top["dH"], s2d=top["s2d_out"])
        for rec in self.layers:
            with torch.cuda.stream(stream_of(rec["l"])):
                ops.fill_zero(rec["dc"])
        for wave in range(T + L - 1):
            for l in range(L - 1, -1, -1):
                k = wave - (L - 1 - l)  # the k-th backward step of layer l: t = T - 1 - k
                if k < 0 or k >= T:
                    continue
                t = T - 1 - k
                rec = self.layers[l]
                st = stream_of(l)
                m = B * rec["ho"] * rec["wo"]
                dH = Act(rec["dH"], T * B, 1, rec["ho"], rec["wo"], he, 0, he)
                with torch.cuda.stream(st):
                    if l < L - 1 and t == T - 1:
                        st.wait_event(self._ev_b[l + 1][t])  # dH[t] of this layer exists
                    ops.clstm_gates_bwd(rec["gact"][t * B:(t + 1) * B].view(m, 4 * he),
                                        rec["c"][(t - 1) * B:t * B] if t > 0 else None, rec["c"][t * B:(t + 1) * B],
                                        self._step(dH, t).buf, rec["dc"], self._step(rec["dpre"], t).buf,
                                        unit_major=rec["um"])
                    if l > 0:  # this step's gradient for the layer below: x-convolution', BN'/pool' of its step t
                        low = self.layers[l - 1]
                        ops.conv3d(self._step(rec["dpre"], t), rec["wx_d"], self._step(rec["g_x"], t), rec["xk"],
                                   (1, 1, 1), rec["xdpf"], plan=self._small_plan)
                        dh_low = low["dH"][t * B:(t + 1) * B]
                        ops.bn_pool2d_bwd(self._step(low["g_pooled"], t).buf, low["argmax"][t * B:(t + 1) * B],
                                          self.bn_scale, dh_low, s2d=low["s2d_out"])
                        self._ev_b[l][t].record(st)
                    if t > 0:
                        if l < L - 1:
                            st.wait_event(self._ev_b[l + 1][t - 1])  # dH[t-1] holds the upper layer's part before we add
                        dprev = self._step(dH, t - 1)
                        ops.conv3d(self._step(rec["dpre"], t), rec["wh_d"], dprev, (1, 5, 5), (1, 1, 1), (0, 2, 2),
                                   acc_in=dprev, plan=rec.get("plan_hd"))
        for l in range(1, L):
            main.wait_stream(self._lanes[l])
        r0 = self.layers[0]
        ops.conv3d(r0["dpre"], r0["wx_d"], r0["g_x"], r0["xk"], (1, 1, 1), r0["xdpf"])

    @_lib.on_device
    def forward(self, mask=None, perturb="reverse"):
        self._mask, self._perturb = (self.zero_mask if mask is None else mask), perturb
        self.generation += 1
        ops.perturb_fwd(self.x, self._mask, perturb, self.in_fmt, self.xin.buf)
        B, T, he = self.B, self.T, self.he
        if self.wave:
            self._forward_wave()
            if self.w_fc is None:
                return None
            ops.head_fwd(self._feat(self.layers[-1]["pooled"], self.eff[-1]), self.w_fc, self.b_fc, self.softmax,
                         self.probs, self.logits, workspace=self.head_ws)
            return self.probs
        for rec in self.layers:
            ops.conv3d(rec["x"], rec["wx_f"], rec["gx"], rec["xk"], rec["xs"], rec["xpf"], scale=rec["ones"],
                       shift=rec["bias"])
            m = B * rec["ho"] * rec["wo"]
            for t in range(T):
                gx_t = self._step(rec["gx"], t)
                c_prev = rec["c"][(t - 1) * B:t * B] if t > 0 else None
                c_next, gact_t = rec["c"][t * B:(t + 1) * B], rec["gact"][t * B:(t + 1) * B].view(m, 4 * he)
                if t > 0 and rec["um"]:  # convolution + gates + state update in one launch
                    ops.conv_lstm_step(self._step(rec["h"], t - 1), rec["wh_f"], gx_t, c_prev, c_next,
                                       self._step(rec["h"], t), gact_t, (1, 5, 5), (0, 2, 2), plan=rec.get("plan_hf"))
                    continue
                if t > 0:
                    ops.conv3d(self._step(rec["h"], t - 1), rec["wh_f"], gx_t, (1, 5, 5), (1, 1, 1), (0, 2, 2), acc_in=gx_t,
                               plan=rec.get("plan_hf"))
                ops.clstm_gates_fwd(gx_t.buf.view(m, 4 * he), c_prev, c_next, self._step(rec["h"], t).buf, gact_t,
                                    unit_major=rec["um"])
            ops.bn_pool2d_fwd(rec["h"].buf.view(T * B, rec["ho"], rec["wo"], he), self.bn_scale, self.bn_shift,
                              rec["pooled"].buf, rec["argmax"], s2d=rec["s2d_out"])
        if self.w_fc is None:
            return None
        te = self.eff[-1]
        ops.head_fwd(self._feat(self.layers[-1]["pooled"], te), self.w_fc, self.b_fc, self.softmax, self.probs,
                     self.logits, workspace=self.head_ws)
        return self.probs

    def step_outputs(self):
        """What ConvLSTM.forward returns (pt/models/convolution_lstm.py:96-132): the last layer's BN+pooled output at
        every effective step as fp32 NCHW tensors, and (x, new_c) = that output and the last layer's cell state at
        the final step."""
        top = self.layers[-1]
        B, T, hid = self.B, self.T, self.hid
        pooled = top["pooled"].buf.view(T, B, self.fh, self.fw, self.he)[..., :hid]
        outs = [pooled[t].permute(0, 3, 1, 2).float().contiguous() for t in self.eff]
        last = pooled[T - 1].permute(0, 3, 1, 2).float().contiguous()
        c = top["c"].view(T, B, top["ho"], top["wo"], self.he)[T - 1, ..., :hid].permute(0, 3, 1, 2).contiguous()
        return outs, (last, c)

    # ------------------------------------------------------------------ backward (BPTT data gradient)
    @_lib.on_device
    def backward(self, to_mask=True):
        B, T, he = self.B, self.T, self.he
        te = self.eff[-1]
        top = self.layers[-1]
        # only the last effective step feeds the classifier: its rows of g_pooled are rewritten, the rest stay 0
        ops.head_bwd(self._feat(top["g_pooled"], te), self.w_fc, self.softmax, self.probs, self.dprobs)
        if self.wave:
            self._wave_setup()
            self._backward_wave()
            if to_mask:
                ops.perturb_bwd(self.x, self._mask, self._perturb, self.in_fmt, self.g_xin.buf, self.dm)
            return self.dm
        for rec in reversed(self.layers):
            ops.bn_pool2d_bwd(rec["g_pooled"].buf, rec["argmax"], self.bn_scale, rec["dH"], s2d=rec["s2d_out"])
            ops.fill_zero(rec["dc"])
            m = B * rec["ho"] * rec["wo"]
            dH = Act(rec["dH"], T * B, 1, rec["ho"], rec["wo"], he, 0, he)
            for t in range(T - 1, -1, -1):
                ops.clstm_gates_bwd(rec["gact"][t * B:(t + 1) * B].view(m, 4 * he),
                                    rec["c"][(t - 1) * B:t * B] if t > 0 else None, rec["c"][t * B:(t + 1) * B],
                                    self._step(dH, t).buf, rec["dc"], self._step(rec["dpre"], t).buf,
                                    unit_major=rec["um"])
                if t > 0:
                    dprev = self._step(dH, t - 1)
                    if self.mode == "fp32":
                        ops.conv3d(self._step(rec["dpre"], t), rec["wh_d"], dprev, (1, 5, 5), (1, 1, 1), (0, 2, 2),
                                   acc_in=dprev, transposed=1)
                    else:
                        ops.conv3d(self._step(rec["dpre"], t), rec["wh_d"], dprev, (1, 5, 5), (1, 1, 1), (0, 2, 2),
                                   acc_in=dprev, plan=rec.get("plan_hd"))
            if self.mode == "fp32":
                ops.conv3d(rec["dpre"], rec["wx_d"], rec["g_x"], (1, 5, 5), (1, 2, 2), (0, 2, 2), transposed=1)
            else:
                ops.conv3d(rec["dpre"], rec["wx_d"], rec["g_x"], rec["xk"], (1, 1, 1), rec["xdpf"])
        if to_mask:
            ops.perturb_bwd(self.x, self._mask, self._perturb, self.in_fmt, self.g_xin.buf, self.dm)
        return self.dm

    # ------------------------------------------------------------------ Grad-CAM operands
    @_lib.on_device
    def gradcam_operands(self):
        """Activations / gradients of the stacked effective-step outputs (pt/pytorch-grad-cam/grad-cam.py:42-49,
        pt/grad_cam_videos.py:87-91): [B, E, h, w, hid] channels-last; only the last step has a gradient
        (the classifier reads output[-1] only)."""
        top = self.layers[-1]
        B, E = self.B, len(self.eff)
        hw, he = self.fh * self.fw, self.he
        pooled = top["pooled"].buf.view(self.T, B, hw * he)
        act = pooled[self.eff].permute(1, 0, 2).contiguous().view(B, E, self.fh, self.fw, he)
        grad = ops.zeros((B, E, hw * he), torch.float32, self.device)
        g = Act(self.g_feat_raw, B, 1, 1, 1, hw * he, 0, hw * he)
        ops.head_bwd(g, self.w_fc, self.softmax, self.probs, self.dprobs)
        grad[:, E - 1] = self.g_feat_raw
        return (Act(act, B, E, self.fh, self.fw, he, 0, he),
                Act(grad.view(B, E, self.fh, self.fw, he), B, E, self.fh, self.fw, he, 0, he))
