"""Python handle on the C-ABI decoder engine (include/b200_decoder.h).

PyTorch is used here for device memory (the flat parameter / gradient / optimizer arenas and the
activation workspace are torch allocations), streams and torch.distributed — nothing else.  All
arithmetic happens in libb200decoder.so; there is no CPU or eager fallback.
"""
import ctypes as C
import math
from typing import Dict, List, Optional, Tuple

import torch

from . import _lib as L

import contextlib
import os


def nvtx_range(name: str):
    """NVTX range around one engine call when B200_NVTX=1 (the C++ engine adds the per-layer / per-part /
    per-position ranges inside).  A no-op context manager otherwise."""
    if os.environ.get("B200_NVTX", "0") not in ("", "0"):
        return torch.cuda.nvtx.range(name)
    return contextlib.nullcontext()


def sinusoid_table(max_len: int, d_model: int) -> torch.Tensor:
    """The `pe` buffer of decoder.PositionalEncodingBatchFirst (reference decoder.py:34-51)."""
    position = torch.arange(max_len).unsqueeze(1)
    div_term = torch.exp(torch.arange(0, d_model, 2) * (-math.log(10000.0) / d_model))
    pe = torch.zeros(max_len, d_model)
    pe[:, 0::2] = torch.sin(position * div_term)
    pe[:, 1::2] = torch.cos(position * div_term)
    return pe.unsqueeze(0)


class DecoderEngine:
    """Owns the flat arenas and the workspace; one instance per model replica (one per GPU)."""

    def __init__(self, vocab_size: int, embed_dim: int, num_heads: int, num_layers: int, ff_dim: int,
                 max_seq_len: int, pad_idx: int = 0, enc_dim: Optional[int] = None,
                 act: str = "relu", device="cuda"):
        self.lib = L.lib()
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("b200 decoder engine needs a CUDA device (sm_100a); there is no CPU fallback")
        L.check(self.lib.b200_check_device(self.device.index or 0), "b200_check_device")
        self.cfg = L.EngineConfig()
        self.cfg.vocab_size, self.cfg.embed_dim, self.cfg.num_heads = vocab_size, embed_dim, num_heads
        self.cfg.num_layers, self.cfg.ff_dim, self.cfg.max_seq_len = num_layers, ff_dim, max_seq_len
        self.cfg.enc_dim = enc_dim if enc_dim is not None else embed_dim
        self.cfg.pad_idx = pad_idx
        self.cfg.ln_eps = 1e-5
        self.cfg.act = {"relu": 1, "gelu": 2}[act]
        self.handle = C.c_void_p()
        L.check(self.lib.b200_engine_create(C.byref(self.cfg), C.byref(self.handle)), "engine_create")
        self.vocab_size, self.embed_dim, self.num_heads = vocab_size, embed_dim, num_heads
        self.num_layers, self.ff_dim, self.max_seq_len = num_layers, ff_dim, max_seq_len
        self.enc_dim = self.cfg.enc_dim
        self.pad_idx = pad_idx

        self.total = int(self.lib.b200_engine_param_count(self.handle))
        self.layout: Dict[str, Tuple[int, int]] = {}
        buf = C.create_string_buffer(256)
        for i in range(self.lib.b200_engine_num_params(self.handle)):
            L.check(self.lib.b200_engine_param_name(self.handle, i, buf, 256), "param_name")
            name = buf.value.decode()
            numel = C.c_int64()
            off = self.lib.b200_engine_param_offset(self.handle, name.encode(), C.byref(numel))
            self.layout[name] = (int(off), int(numel.value))

        with torch.cuda.device(self.device):
            self.params = torch.zeros(self.total, device=self.device, dtype=torch.float32)
            self.params_bf16 = torch.zeros(self.total, device=self.device, dtype=torch.bfloat16)
            self.grads = torch.zeros(self.total, device=self.device, dtype=torch.float32)
            self.pe = sinusoid_table(max_seq_len, embed_dim).to(self.device).contiguous()
            self._scal = torch.zeros(8, device=self.device, dtype=torch.float32)
            self._step_dev = torch.zeros(1, device=self.device, dtype=torch.int32)
            self._lr_dev = torch.zeros(1, device=self.device, dtype=torch.float32)
        self._lr_host = None
        self.exp_avg = None
        self.exp_avg_sq = None
        self.opt_step = 0
        self._ws = None
        self._ws_key = None
        self._shadow_version = -1
        self._shadow_fresh = False      # True while the last writer of the master weights was the fused AdamW kernel
        self._graph_pins = set()        # captured CUDA graphs holding raw pointers into the workspace / dropout state
        self._scratch_grads = None
        L.check(self.lib.b200_engine_bind(self.handle, L.ptr(self.params), L.ptr(self.params_bf16),
                                          L.ptr(self.grads), L.ptr(self.pe)), "engine_bind")

    def __del__(self):
        try:
            if getattr(self, "handle", None):
                self.lib.b200_engine_destroy(self.handle)
                self.handle = None
        except Exception:
            pass

    # ------------------------------------------------------------------ dropout
    def set_dropout(self, p: float, seed: int = 0) -> None:
        """Enable dropout (probability p at the reference's 1 + 6 L sites) for forwards with
        training=True; p = 0 disables.  Masks are a pure function of (seed, step counter, site,
        element), so the device state [seed, counter] is all that is kept."""
        p = float(p)
        if p > 0:
            if getattr(self, "_graph_pins", None) and getattr(self, "_drop_state", None) is not None:
                # a captured graph reads this buffer: re-seed in place instead of reallocating it
                self._drop_state.copy_(torch.tensor([seed & 0x7FFFFFFF, 0], dtype=torch.int32))
            else:
                self._drop_state = torch.tensor([seed & 0x7FFFFFFF, 0], device=self.device, dtype=torch.int32)
        else:
            if getattr(self, "_graph_pins", None) and getattr(self, "dropout_p", 0.0) > 0:
                raise RuntimeError("b200 engine: a captured CUDA graph uses the dropout state; release it "
                                   "(GraphedTrainStep.release()) before disabling dropout")
            self._drop_state = None
        self.dropout_p = p
        self._drop_active = None
        self.dropout_active(True)

    def dropout_active(self, active: bool) -> None:
        """Train-mode switch (nn.Module.train()/eval()): dropout applies only while active."""
        active = bool(active) and getattr(self, "dropout_p", 0.0) > 0
        if getattr(self, "_drop_active", None) is active:
            return
        self._drop_active = active
        L.check(self.lib.b200_engine_set_dropout(self.handle, self.dropout_p if active else 0.0,
                                                 L.ptr(self._drop_state) if active else None), "set_dropout")

    def dropout_state(self):
        """(seed, counter) as host ints; the counter is the value the LAST training forward used."""
        if getattr(self, "_drop_state", None) is None:
            return None
        s = self._drop_state.cpu()
        return int(s[0]), int(s[1])

    # ------------------------------------------------------------------ arenas
    def shapes(self) -> Dict[str, Tuple[int, ...]]:
        E, F, V = self.embed_dim, self.ff_dim, self.vocab_size
        sh: Dict[str, Tuple[int, ...]] = {}
        if self.enc_dim != E:
            sh["projection.weight"] = (E, self.enc_dim)
            sh["projection.bias"] = (E,)
        sh["token_embedding.weight"] = (V, E)
        for l in range(self.num_layers):
            p = f"transformer_decoder.layers.{l}."
            sh[p + "self_attn.in_proj_weight"] = (3 * E, E)
            sh[p + "self_attn.in_proj_bias"] = (3 * E,)
            sh[p + "self_attn.out_proj.weight"] = (E, E)
            sh[p + "self_attn.out_proj.bias"] = (E,)
            sh[p + "multihead_attn.in_proj_weight"] = (3 * E, E)
            sh[p + "multihead_attn.in_proj_bias"] = (3 * E,)
            sh[p + "multihead_attn.out_proj.weight"] = (E, E)
            sh[p + "multihead_attn.out_proj.bias"] = (E,)
            sh[p + "linear1.weight"] = (F, E)
            sh[p + "linear1.bias"] = (F,)
            sh[p + "linear2.weight"] = (E, F)
            sh[p + "linear2.bias"] = (E,)
            for n in ("norm1", "norm2", "norm3"):
                sh[p + n + ".weight"] = (E,)
                sh[p + n + ".bias"] = (E,)
        sh["fc_out.weight"] = (V, E)
        sh["fc_out.bias"] = (V,)
        return sh

    def view(self, name: str, arena: Optional[torch.Tensor] = None) -> torch.Tensor:
        off, n = self.layout[name]
        a = self.params if arena is None else arena
        return a[off:off + n].view(self.shapes()[name])

    def load(self, tensors: Dict[str, torch.Tensor]) -> None:
        """Copy named fp32 tensors (reference state_dict names relative to the decoder) in."""
        with torch.no_grad():
            for k, v in tensors.items():
                if k in self.layout:
                    self.view(k).copy_(v.to(self.device, torch.float32))
        self.mark_weights_changed()

    def sync_shadow(self, force: bool = False) -> None:
        """Refresh the bf16 weight shadow (what every GEMM reads) from the fp32 master.

        In-place edits of a parameter tensor itself bump the arena's version counter and are detected;
        edits through `p.data` (p.data.copy_/mul_, EMA swaps, dist.broadcast(p.data), HF-style init
        utilities) are invisible to any counter.  Policy: the shadow is trusted only while the last
        writer of the master was the fused AdamW kernel (`_shadow_fresh`, which writes master and
        shadow together) -- i.e. inside the fused training loop; every other entry point
        (eval-mode forwards, generation, train()/eval() switches, load) re-casts, one pass over the
        arena.  After editing `p.data` in the middle of a fused training loop, call
        `engine.sync_shadow(force=True)`."""
        ver = self.params._version
        if force or not self._shadow_fresh or ver != self._shadow_version:
            L.check(self.lib.b200_cast_f32_to_bf16(L.ptr(self.params), L.ptr(self.params_bf16),
                                                   C.c_int64(self.total), L.cur_stream()), "cast")
            self._shadow_version = ver

    def mark_weights_changed(self) -> None:
        """The fp32 master may have been edited behind the engine's back: re-cast now and stop trusting the shadow."""
        self._shadow_fresh = False
        self.sync_shadow(force=True)

    # ------------------------------------------------------------------ CUDA-graph pins
    def pin_for_graph(self, owner) -> None:
        self._graph_pins.add(id(owner))

    def unpin_for_graph(self, owner) -> None:
        self._graph_pins.discard(id(owner))

    def reserve_workspace(self, B: int, T: int, S: int, mem_dim: int, training: bool) -> None:
        """Grow the activation workspace to fit this shape now (call before capturing a CUDA graph
        when a larger evaluation / logits shape will run between replays)."""
        self._ensure_ws(B, T, S, mem_dim, training)

    def zero_grad(self) -> None:
        self.grads.zero_()

    def _bind(self, grads: torch.Tensor) -> None:
        L.check(self.lib.b200_engine_bind(self.handle, L.ptr(self.params), L.ptr(self.params_bf16),
                                          L.ptr(grads), L.ptr(self.pe)), "engine_bind")

    def scratch_grads(self):
        """Context manager: the engine's backward accumulates into a zeroed scratch arena instead of
        `self.grads` (the autograd bridge returns per-parameter views of it, so whatever the fused
        path accumulated in `self.grads` is neither cloned nor disturbed).  The scratch arena is
        reused by the next call: tensors obtained through torch.autograd.grad alias it until then."""
        import contextlib

        @contextlib.contextmanager
        def cm():
            if self._scratch_grads is None:
                self._scratch_grads = torch.zeros_like(self.grads)
            else:
                self._scratch_grads.zero_()
            self._bind(self._scratch_grads)
            try:
                yield self._scratch_grads
            finally:
                self._bind(self.grads)
        return cm()

    # ------------------------------------------------------------------ workspace
    def _ensure_ws(self, B: int, T: int, S: int, mem_dim: int, training: bool) -> None:
        need = int(self.lib.b200_engine_workspace_bytes(self.handle, B, T, S, mem_dim, int(training)))
        if self._ws is None or self._ws.numel() < need:
            if self._graph_pins:
                raise RuntimeError(
                    "b200 engine: this forward needs a %d-byte workspace but a captured CUDA graph holds pointers into "
                    "the current %d-byte one; call engine.reserve_workspace(B, T, S, mem_dim, training) for the largest "
                    "shape before capturing, or GraphedTrainStep.release() first" % (need, 0 if self._ws is None else self._ws.numel()))
            self._ws = None
            self._ws = torch.empty(need, device=self.device, dtype=torch.uint8)
            L.check(self.lib.b200_engine_set_workspace(self.handle, L.ptr(self._ws), C.c_int64(self._ws.numel())),
                    "set_workspace")

    def _mem(self, memory: torch.Tensor) -> torch.Tensor:
        """fp32 features are cast by the engine; bf16 features (e.g. model.FeatureCache) are consumed in place."""
        is_bf16 = memory.dtype == torch.bfloat16
        memory = memory.contiguous() if is_bf16 else memory.to(torch.float32).contiguous()
        if getattr(self, "_mem_bf16", None) is not is_bf16:
            L.check(self.lib.b200_engine_set_memory_dtype(self.handle, int(is_bf16)), "set_memory_dtype")
            self._mem_bf16 = is_bf16
        return memory

    def _prep(self, tokens, memory, mem_pad):
        assert tokens.dtype == torch.int64 and tokens.is_cuda and tokens.dim() == 2
        tokens = tokens.contiguous()
        memory = self._mem(memory)
        assert memory.dim() == 3 and memory.shape[0] == tokens.shape[0]
        if mem_pad is not None:
            mem_pad = mem_pad.to(torch.uint8).contiguous()
        return tokens, memory, mem_pad

    # ------------------------------------------------------------------ whole-model calls
    def forward_logits(self, tokens, memory, mem_pad=None, training: bool = False) -> torch.Tensor:
        tokens, memory, mem_pad = self._prep(tokens, memory, mem_pad)
        B, T = tokens.shape
        S, mem_dim = memory.shape[1], memory.shape[2]
        self.sync_shadow()
        self._ensure_ws(B, T, S, mem_dim, training)
        logits = torch.empty(B, T, self.vocab_size, device=self.device, dtype=torch.float32)
        self._keep = (tokens, memory, mem_pad)
        with nvtx_range("b200.forward_logits"):
            L.check(self.lib.b200_engine_forward_logits(self.handle, L.ptr(tokens), L.ptr(memory), L.ptr(mem_pad),
                                                        B, T, S, mem_dim, int(training), L.ptr(logits),
                                                        L.cur_stream()), "forward_logits")
        return logits

    def forward_loss(self, tokens, targets, memory, mem_pad=None, ignore_index: int = 0,
                     training: bool = False, lengths=None) -> torch.Tensor:
        """Returns a device tensor [mean_loss, n_valid_targets].

        lengths (optional, host list / CPU tensor of B ints): PACKED (var-len) forward -- sample b keeps only its first
        lengths[b] positions (its non-PAD prefix: the reference pads every caption to MAX_SEQ_LEN, tokenizer.py:293-313,
        dataset.py:176-206) and every row-wise kernel runs on sum(lengths) rows instead of B * T.  Loss and gradients
        equal the padded call's.  `packed_lengths()` derives the lengths from the token ids."""
        tokens, memory, mem_pad = self._prep(tokens, memory, mem_pad)
        targets = targets.contiguous()
        B, T = tokens.shape
        S, mem_dim = memory.shape[1], memory.shape[2]
        self.sync_shadow()
        self._ensure_ws(B, T, S, mem_dim, training)
        out = torch.empty(2, device=self.device, dtype=torch.float32)
        self._keep = (tokens, memory, mem_pad, targets)
        if lengths is not None:
            lens = torch.as_tensor(lengths, dtype=torch.int32, device="cpu").reshape(-1)
            if lens.numel() != B or int(lens.min()) < 1 or int(lens.max()) > T:
                raise ValueError("forward_loss: lengths must hold B values in [1, T]")
            # row offsets travel through a small ring of pinned staging buffers: a pageable H2D copy would block the
            # host until the stream drains and expose the launch time of the whole eager step
            ring = getattr(self, "_cu_ring", None)
            if ring is None or ring[0][0].numel() < B + 1:
                ring = [(torch.empty(B + 1, dtype=torch.int32).pin_memory(),
                         torch.empty(B + 1, dtype=torch.int32, device=self.device), torch.cuda.Event()) for _ in range(4)]
                self._cu_ring, self._cu_next = ring, 0
            cu, cu_dev, ev = ring[self._cu_next]
            self._cu_next = (self._cu_next + 1) % len(ring)
            ev.synchronize()                      # the copy that last used this staging buffer has finished
            cu[0] = 0
            torch.cumsum(lens, 0, out=cu[1:B + 1])
            total = int(cu[B])
            cu_dev[:B + 1].copy_(cu[:B + 1], non_blocking=True)
            ev.record()
            self._keep = self._keep + (cu_dev,)
            with nvtx_range("b200.forward_loss_packed"):
                L.check(self.lib.b200_engine_forward_loss_packed(
                    self.handle, L.ptr(tokens), L.ptr(targets), L.ptr(memory), L.ptr(mem_pad), B, T, S, mem_dim,
                    C.c_int64(ignore_index), int(training), L.ptr(cu_dev), total, L.ptr(out), L.cur_stream()),
                    "forward_loss_packed")
            return out
        with nvtx_range("b200.forward_loss"):
            L.check(self.lib.b200_engine_forward_loss(self.handle, L.ptr(tokens), L.ptr(targets), L.ptr(memory),
                                                      L.ptr(mem_pad), B, T, S, mem_dim, C.c_int64(ignore_index),
                                                      int(training), L.ptr(out), L.cur_stream()), "forward_loss")
        return out

    @staticmethod
    def packed_lengths(tokens: torch.Tensor, pad_idx: int = 0):
        """Per-sample length of the non-PAD prefix of [B, T] token ids (on the host), or None when some caption has a
        PAD *inside* it (the packed path keeps prefixes only; such a batch has to take the padded path, where PAD keys
        are masked like the reference does, decoder.py:158-162)."""
        nonpad = (tokens != pad_idx)
        prefix = nonpad.long().cumprod(1).sum(1)
        if not torch.equal(prefix, nonpad.sum(1)) or int(prefix.min()) < 1:
            return None
        return prefix.to("cpu", torch.int32)

    def backward(self, inv_count: Optional[torch.Tensor] = None, want_dmemory: bool = False,
                 events: Optional[List[torch.cuda.Event]] = None) -> Optional[torch.Tensor]:
        dmem = None
        if want_dmemory:
            dmem = torch.empty_like(self._keep[1], dtype=torch.float32)
        ev_arr, n_ev = None, 0
        if events:
            n_ev = len(events)
            ev_arr = (C.c_void_p * n_ev)(*[C.c_void_p(e.cuda_event) for e in events])
        with nvtx_range("b200.backward"):
            L.check(self.lib.b200_engine_backward(self.handle, L.ptr(inv_count), L.ptr(dmem), ev_arr, n_ev,
                                                  L.cur_stream()), "backward")
        return dmem

    def backward_parts(self, first: int, last: int, inv_count: Optional[torch.Tensor] = None) -> None:
        """Parts [first, last] of backward (0 = LM head, k = layer L-k, L+1 = embedding/projection)."""
        with nvtx_range("b200.backward_part"):
            L.check(self.lib.b200_engine_backward_parts(self.handle, L.ptr(inv_count), None, first, last, L.cur_stream()),
                    "backward_parts")

    def backward_from_dlogits(self, dlogits: torch.Tensor, want_dmemory: bool = False) -> Optional[torch.Tensor]:
        dlogits = dlogits.to(torch.float32).contiguous()
        dmem = None
        if want_dmemory:
            dmem = torch.empty_like(self._keep[1], dtype=torch.float32)
        with nvtx_range("b200.backward_from_dlogits"):
            L.check(self.lib.b200_engine_backward_from_dlogits(self.handle, L.ptr(dlogits), L.ptr(dmem),
                                                               L.cur_stream()), "backward_from_dlogits")
        return dmem

    def grad_buckets(self) -> List[Tuple[int, int]]:
        n = self.lib.b200_engine_grad_buckets(self.handle, None, None, 0)
        offs = (C.c_int64 * n)()
        cnts = (C.c_int64 * n)()
        self.lib.b200_engine_grad_buckets(self.handle, offs, cnts, n)
        return [(int(offs[i]), int(cnts[i])) for i in range(n)]

    # ------------------------------------------------------------------ optimizer
    def norm_begin(self) -> None:
        """Start a per-bucket gradient-norm accumulation (see norm_add_bucket): zero the squared-norm scalar."""
        self._scal[0:1].zero_()

    def norm_add_bucket(self, index: int) -> None:
        """sumsq += |g|^2 over gradient bucket `index` (readiness order of grad_buckets()) on the CURRENT stream.  Called
        on a side stream behind the bucket's event, the global-norm pass of clip_grad_norm_ (train.py:97) overlaps the rest
        of backward and reads every bucket while it is still in L2; adamw_step(norm_ready=True) then skips its own pass."""
        off, cnt = self.grad_buckets()[index]
        L.check(self.lib.b200_grad_sumsq(C.c_void_p(self.grads.data_ptr() + 4 * off), C.c_int64(cnt), L.ptr(self._scal[0:1]),
                                         L.cur_stream()), "grad_sumsq")

    def buckets_cover_arena(self) -> bool:
        b = sorted(self.grad_buckets())
        pos = 0
        for off, cnt in b:
            if off != pos:
                return False
            pos = off + cnt
        return pos == self.total

    def adamw_step(self, lr=1e-4, betas=(0.9, 0.98), eps=1e-9, weight_decay=1e-5, max_norm=5.0,
                   norm_ready: bool = False) -> torch.Tensor:
        """clip_grad_norm_(max_norm) + AdamW over the whole arena (train.py:96-100, 319-325).
        Returns the device scalar holding the squared global gradient norm (pre-clip)."""
        if self.exp_avg is None:
            self.exp_avg = torch.zeros_like(self.params)
            self.exp_avg_sq = torch.zeros_like(self.params)
        self.opt_step += 1           # host mirror of the device-side counter (state_dict / resume)
        if self._lr_host != lr and not torch.cuda.is_current_stream_capturing():
            self._lr_dev.fill_(lr)   # lr lives on the device so that a captured step can be replayed
            self._lr_host = lr
        sumsq = self._scal[0:1]
        st = L.cur_stream()
        if not norm_ready:
            sumsq.zero_()
            with nvtx_range("b200.optimizer.grad_norm"):
                L.check(self.lib.b200_grad_sumsq(L.ptr(self.grads), C.c_int64(self.total), L.ptr(sumsq), st), "grad_sumsq")
        with nvtx_range("b200.optimizer.clip_adamw"):
            L.check(self.lib.b200_adamw_step_dev(L.ptr(self.params), L.ptr(self.params_bf16), L.ptr(self.grads),
                                                 L.ptr(self.exp_avg), L.ptr(self.exp_avg_sq), C.c_int64(self.total),
                                                 L.ptr(sumsq), C.c_float(max_norm), L.ptr(self._lr_dev),
                                                 C.c_float(betas[0]), C.c_float(betas[1]), C.c_float(eps),
                                                 C.c_float(weight_decay), L.ptr(self._step_dev), st), "adamw_step")
        # the kernel wrote both the fp32 master and the bf16 shadow: nothing to re-sync
        self._shadow_fresh = True
        return sumsq

    # ------------------------------------------------------------------ KV-cached generation
    def decode_begin(self, memory: torch.Tensor, mem_pad: Optional[torch.Tensor] = None, beam: int = 1,
                     max_len: int = 100) -> None:
        """Project the image memory once per image into per-layer cross K/V and reset the caches."""
        memory = self._mem(memory.to(self.device))
        if mem_pad is not None:
            mem_pad = mem_pad.to(self.device, torch.uint8).contiguous()
        B, S, mem_dim = memory.shape
        self.sync_shadow()
        need = int(self.lib.b200_engine_decode_workspace_bytes(self.handle, B, beam, S, mem_dim, max_len))
        if need < 0:
            raise RuntimeError("decode_begin: invalid shape")
        if getattr(self, "_dws", None) is None or self._dws.numel() < need:
            self._dws = None
            self._dws = torch.empty(need, device=self.device, dtype=torch.uint8)
        self._dkeep = (memory, mem_pad)
        self._dshape = (B, beam, max_len)
        with nvtx_range("b200.decode_begin"):
            L.check(self.lib.b200_engine_decode_begin(self.handle, L.ptr(memory), L.ptr(mem_pad), B, beam, S, mem_dim,
                                                      max_len, L.ptr(self._dws), self._dws.numel(), L.cur_stream()),
                    "decode_begin")

    def decode_plan_info(self) -> dict:
        """How the current decode plan is scheduled (after decode_begin); reporting only."""
        import ctypes
        buf = (ctypes.c_int32 * 6)()
        L.check(self.lib.b200_engine_decode_plan_info(self.handle, ctypes.addressof(buf), 6), "decode_plan_info")
        keys = ("partitions", "attention_sms", "split_k_e", "split_k_f", "gemm_grid_cap", "hint_flags")
        return dict(zip(keys, (int(v) for v in buf)))

    def decode_step(self, tokens_in: torch.Tensor, pos: int) -> torch.Tensor:
        tokens_in = tokens_in.to(self.device, torch.int64).contiguous()
        out = torch.empty_like(tokens_in)
        L.check(self.lib.b200_engine_decode_step(self.handle, L.ptr(tokens_in), pos, L.ptr(out), L.cur_stream()),
                "decode_step")
        return out

    def generate_greedy(self, start_id: int, end_id: int, max_len: int, stop_check_interval: int = 0):
        """Returns (tokens [B,max_len] int64: START .. END then PAD, lengths [B] int32) on device."""
        B, beam, plan_len = self._dshape
        assert beam == 1
        toks = torch.empty(B, max_len, device=self.device, dtype=torch.int64)
        lens = torch.empty(B, device=self.device, dtype=torch.int32)
        with nvtx_range("b200.generate_greedy"):
            L.check(self.lib.b200_engine_generate_greedy(self.handle, start_id, end_id, max_len, stop_check_interval,
                                                         L.ptr(toks), L.ptr(lens), L.cur_stream()), "generate_greedy")
        return toks, lens

    def generate_beam(self, start_id: int, end_id: int, max_len: int):
        """Returns (tokens [B,plan max_len], lengths [B], scores [B]) of the best hypothesis."""
        B, beam, plan_len = self._dshape
        toks = torch.empty(B, plan_len, device=self.device, dtype=torch.int64)
        lens = torch.empty(B, device=self.device, dtype=torch.int32)
        score = torch.empty(B, device=self.device, dtype=torch.float32)
        with nvtx_range("b200.generate_beam"):
            L.check(self.lib.b200_engine_generate_beam(self.handle, start_id, end_id, max_len, L.ptr(toks), L.ptr(lens),
                                                       L.ptr(score), L.cur_stream()), "generate_beam")
        return toks, lens, score
