"""Drop-in for the reference's decoder.py: same classes, constructor arguments, forward contract and
state_dict keys (reference decoder.py:16-72, 75-193), with the arithmetic done by the sm_100a engine.

    TransformerDecoder(vocab_size, embed_dim, num_heads, num_layers, ff_dim, max_seq_len,
                       dropout=0.1, pad_idx=0)
    .forward(tgt_tokens (B,T) int64, memory (B,S,E), memory_padding_mask=None (B,S) bool)
        -> logits (B,T,V) fp32

Parameters are nn.Parameters that are VIEWS into the engine's flat fp32 arena, laid out under the
module names torch.nn.TransformerDecoder would produce, so reference checkpoints load with
load_state_dict(strict=True) and ours load in the reference.  Gradients of the fused train step
land in a second flat arena whose views are exposed as .grad.
"""
import math
import warnings
from typing import Optional

import torch
import torch.nn as nn

from .engine import DecoderEngine, sinusoid_table


class PositionalEncodingBatchFirst(nn.Module):
    """x + pe[:, :T] then dropout, batch-first (reference decoder.py:16-72).  Kept as a module so
    that `positional_encoding.pe` stays a state_dict key; inside TransformerDecoder.forward the
    addition is fused into the embedding kernel."""

    def __init__(self, d_model: int, dropout: float = 0.1, max_len: int = 5000, pe: Optional[torch.Tensor] = None):
        super().__init__()
        self.dropout = nn.Dropout(p=dropout)
        self.register_buffer("pe", sinusoid_table(max_len, d_model) if pe is None else pe)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        return self.dropout(x + self.pe[:, :x.size(1), :])


class _Holder(nn.Module):
    """Parameter container giving arena views the reference's attribute paths."""


def _reference_init(vocab_size, embed_dim, num_heads, num_layers, ff_dim, pad_idx):
    """Values the reference constructor would produce for the current torch RNG state: builds the
    same torch.nn modules in the same order as reference decoder.py:105-124 and applies the
    Xavier pass of decoder.py:128-132 (which also overwrites the embedding's padding row)."""
    emb = nn.Embedding(vocab_size, embed_dim, padding_idx=pad_idx)
    layer = nn.TransformerDecoderLayer(d_model=embed_dim, nhead=num_heads, dim_feedforward=ff_dim,
                                       dropout=0.0, batch_first=True)
    dec = nn.TransformerDecoder(layer, num_layers=num_layers)
    fc = nn.Linear(embed_dim, vocab_size)
    for q in list(emb.parameters()) + list(dec.parameters()) + list(fc.parameters()):
        if q.dim() > 1:
            nn.init.xavier_uniform_(q)
    out = {"token_embedding.weight": emb.weight.detach()}
    for k, v in dec.state_dict().items():
        out["transformer_decoder." + k] = v.detach()
    out["fc_out.weight"], out["fc_out.bias"] = fc.weight.detach(), fc.bias.detach()
    return out


class _DecoderFunction(torch.autograd.Function):
    """Autograd bridge for callers that follow the reference loop (logits -> criterion ->
    loss.backward(), train.py:83-93).  Backward consumes dlogits and returns per-parameter
    gradients computed by the engine."""

    @staticmethod
    def forward(ctx, module, tokens, memory, mem_pad, *params):
        eng = module.engine
        logits = eng.forward_logits(tokens, memory, mem_pad, training=True)
        eng._generation = getattr(eng, "_generation", 0) + 1
        ctx.module = module
        ctx.generation = eng._generation
        ctx.memory_needs_grad = memory.requires_grad
        ctx.mem_dim = memory.shape[-1]
        ctx.n_params = len(params)
        return logits

    @staticmethod
    def backward(ctx, dlogits):
        module = ctx.module
        eng = module.engine
        if eng._generation != ctx.generation:
            raise RuntimeError("b200 decoder: the activations of this forward were overwritten by a later "
                               "forward; only one graph per decoder can be alive at a time")
        want_dmem = ctx.memory_needs_grad and ctx.mem_dim == eng.embed_dim
        with eng.scratch_grads() as sg:    # whatever the fused path accumulated in eng.grads stays untouched
            dmem = eng.backward_from_dlogits(dlogits, want_dmemory=want_dmem)
        grads = tuple(eng.view(name, sg) for name in module._param_names)
        return (None, None, dmem, None) + grads


class TransformerDecoder(nn.Module):
    """Embedding*sqrt(E) + sinusoidal PE -> N post-LN decoder layers (causal self-attention with key
    padding from the token ids, cross-attention over the image memory, ReLU FFN) -> vocabulary
    projection; same mathematics as reference decoder.py:134-193."""

    def __init__(self, vocab_size: int, embed_dim: int, num_heads: int, num_layers: int, ff_dim: int,
                 max_seq_len: int, dropout: float = 0.1, pad_idx: int = 0, *, enc_dim: Optional[int] = None,
                 device=None):
        super().__init__()
        self.embed_dim = embed_dim
        self.pad_idx = pad_idx
        self.vocab_size = vocab_size
        self.num_heads = num_heads
        self.dropout_p = float(dropout)
        dev = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
        init = _reference_init(vocab_size, embed_dim, num_heads, num_layers, ff_dim, pad_idx)
        self.engine = DecoderEngine(vocab_size, embed_dim, num_heads, num_layers, ff_dim, max_seq_len,
                                    pad_idx=pad_idx, enc_dim=enc_dim, device=dev)
        eng = self.engine
        self._param_names = [n for n in eng.layout if not n.startswith("projection.")]

        def P(name):
            p = nn.Parameter(eng.view(name))
            p.grad = eng.view(name, eng.grads)
            return p

        self.token_embedding = _Holder()
        self.token_embedding.weight = P("token_embedding.weight")
        self.positional_encoding = PositionalEncodingBatchFirst(embed_dim, dropout, max_seq_len, pe=eng.pe)
        self.transformer_decoder = _Holder()
        layers = []
        for l in range(num_layers):
            pre = f"transformer_decoder.layers.{l}."
            lay = _Holder()
            for attn in ("self_attn", "multihead_attn"):
                a = _Holder()
                a.in_proj_weight = P(pre + attn + ".in_proj_weight")
                a.in_proj_bias = P(pre + attn + ".in_proj_bias")
                a.out_proj = _Holder()
                a.out_proj.weight = P(pre + attn + ".out_proj.weight")
                a.out_proj.bias = P(pre + attn + ".out_proj.bias")
                setattr(lay, attn, a)
            for lin in ("linear1", "linear2", "norm1", "norm2", "norm3"):
                h = _Holder()
                h.weight = P(pre + lin + ".weight")
                h.bias = P(pre + lin + ".bias")
                setattr(lay, lin, h)
            layers.append(lay)
        self.transformer_decoder.layers = nn.ModuleList(layers)
        self.fc_out = _Holder()
        self.fc_out.weight = P("fc_out.weight")
        self.fc_out.bias = P("fc_out.bias")
        eng.load(init)
        # dropout at the reference's sites (decoder.py:72,112-118) in train mode; the mask generator is
        # counter-based, seeded from torch's CPU generator so torch.manual_seed() controls it
        eng.set_dropout(self.dropout_p, torch.initial_seed() & 0x7FFFFFFF)

    # ------------------------------------------------------------------ nn.Module plumbing
    def _apply(self, fn, recurse=True):
        probe = fn(torch.empty(0, device=self.engine.device))
        if probe.device != self.engine.device or probe.dtype != torch.float32:
            raise RuntimeError("b200 TransformerDecoder lives on its CUDA device in fp32 master / bf16 compute; "
                               "moving or casting it is not supported (no CPU fallback)")
        return self

    def train(self, mode: bool = True):
        """Mode switches are where foreign code typically edits weights through `p.data` (EMA swaps, re-inits):
        the bf16 shadow is re-cast and no longer trusted until the next fused optimizer step (engine.sync_shadow)."""
        if mode != self.training:
            self.engine._shadow_fresh = False
        return super().train(mode)

    def _flat_params(self):
        return [self.get_parameter(n) for n in self._param_names]

    def relink_grads(self) -> None:
        """Point every .grad back at the flat gradient arena (after zero_grad(set_to_none=True))."""
        for n in self._param_names:
            self.get_parameter(n).grad = self.engine.view(n, self.engine.grads)

    def _mask(self, memory_padding_mask):
        return None if memory_padding_mask is None else memory_padding_mask.to(self.engine.device)

    def set_dropout_seed(self, seed: int) -> None:
        """Re-seed the dropout masks (data-parallel replicas use seed + rank)."""
        self.engine.set_dropout(self.dropout_p, seed)

    # ------------------------------------------------------------------ reference API
    def forward(self, tgt_tokens: torch.Tensor, memory: torch.Tensor,
                memory_padding_mask: Optional[torch.Tensor] = None) -> torch.Tensor:
        self.engine.dropout_active(self.training)
        dev = self.engine.device
        tgt_tokens = tgt_tokens.to(dev)
        memory = memory.to(dev)
        mem_pad = self._mask(memory_padding_mask)
        needs_graph = torch.is_grad_enabled() and (memory.requires_grad or
                                                   any(p.requires_grad for p in self._flat_params()))
        if needs_graph:
            return _DecoderFunction.apply(self, tgt_tokens, memory, mem_pad, *self._flat_params())
        return self.engine.forward_logits(tgt_tokens, memory, mem_pad, training=False)

    # ------------------------------------------------------------------ fused fast path
    def loss(self, tgt_tokens, target_tokens, memory, memory_padding_mask=None, ignore_index: int = 0,
             training: Optional[bool] = None, lengths=None) -> torch.Tensor:
        """mean CE over targets != ignore_index with the LM head fused into the loss (logits are
        never materialised).  Returns a device tensor [loss, n_valid]; no host sync.

        lengths (host, B ints; e.g. DecoderEngine.packed_lengths(cpu_tokens)): packed / var-len path -- only the non-PAD
        prefix of every caption is computed (the reference pads to MAX_SEQ_LEN, tokenizer.py:293-313); same loss and
        gradients as the padded call."""
        training = self.training if training is None else training
        self.engine.dropout_active(self.training and training)
        return self.engine.forward_loss(tgt_tokens, target_tokens, memory, self._mask(memory_padding_mask),
                                        ignore_index, training=training, lengths=lengths)

    def backward(self, inv_count=None, events=None):
        """Backward of the last loss(training=True) into the flat gradient arena (accumulates)."""
        return self.engine.backward(inv_count=inv_count, events=events)
