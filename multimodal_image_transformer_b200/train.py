"""Drop-in for the train / eval steps of the reference's train.py (reference train.py:62-151):
same function names, argument lists and return values.  With a B200 model + B200AdamW the step is
the fused path (forward + LM-head/CE without logits + backward + bucketed gradient all-reduce +
global-norm clip + AdamW, no per-step host sync except the loss read the reference also does);
with any other optimizer / criterion it degrades to the reference's own loop over the
autograd-compatible forward."""
from typing import Optional

import os

import torch

from . import config
from .dp import DataParallel


class B200AdamW(torch.optim.Optimizer):
    """torch.optim.AdamW-compatible facade over the engine's fused clip + AdamW kernel
    (reference train.py:96-100, 319-325; torch AdamW semantics).

    It IS a torch.optim.Optimizer (param_groups / defaults / initial_lr), so the reference's optional
    warm-up schedule -- transformers.get_linear_schedule_with_warmup, a LambdaLR, reference
    train.py:331-341 -- drives it: the scheduler writes param_groups[0]["lr"], step() hands that
    value to the device-resident learning rate the (possibly CUDA-graph-replayed) kernel reads."""

    def __init__(self, model, lr=config.LEARNING_RATE, betas=(config.ADAM_BETA1, config.ADAM_BETA2),
                 eps=config.ADAM_EPS, weight_decay=config.WEIGHT_DECAY, max_grad_norm: float = 0.0):
        self.model_ref = model
        self.decoder = model.decoder if hasattr(model, "decoder") else model
        self.engine = self.decoder.engine
        params = [p for p in model.parameters() if p.requires_grad]
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay))
        self.max_grad_norm = max_grad_norm
        self.last_grad_sumsq: Optional[torch.Tensor] = None

    def zero_grad(self, set_to_none: bool = False) -> None:
        self.engine.zero_grad()          # the arena stays allocated; .grad views keep pointing at it

    def step(self, closure=None, max_grad_norm: Optional[float] = None, norm_ready: bool = False) -> None:
        if closure is not None and not callable(closure):      # step(5.0): positional clip value (round-1 signature)
            max_grad_norm, closure = float(closure), None
        if closure is not None:
            raise RuntimeError("B200AdamW.step does not take a closure (the fused step runs forward and backward itself)")
        g = self.param_groups[0]
        mn = self.max_grad_norm if max_grad_norm is None else max_grad_norm
        self.last_grad_sumsq = self.engine.adamw_step(lr=float(g["lr"]), betas=g["betas"], eps=g["eps"],
                                                      weight_decay=g["weight_decay"], max_norm=mn, norm_ready=norm_ready)

    # ---- checkpoint compatibility (reference train.py:351-357, 424): torch.optim.AdamW layout
    def _named_trainable(self):
        """(index in the reference optimizer's parameter list, engine tensor name) of every trainable
        parameter.  The reference builds AdamW over model.parameters() (train.py:319), which lists the
        frozen encoder's parameters first: they keep their slots but never get state."""
        model = self.model_ref
        names, offset = [], 0
        if hasattr(model, "encoder") and hasattr(model, "decoder"):
            offset = sum(1 for _ in model.encoder.parameters())
            if any(n.startswith("projection.") for n in self.engine.layout):
                names += ["projection.weight", "projection.bias"]
        names += [n for n in self.engine.layout if not n.startswith("projection.")]
        return [(offset + i, n) for i, n in enumerate(names)], offset

    def state_dict(self, torch_format: bool = True):
        """torch.optim.AdamW-shaped state ({'state': {idx: {step, exp_avg, exp_avg_sq}}, 'param_groups'}) so a
        checkpoint written here resumes in the reference and vice versa; torch_format=False returns
        the flat-arena form."""
        e = self.engine
        groups = [{k: (float(v) if k == "lr" else v) for k, v in g.items() if k != "params"} for g in self.param_groups]
        if not torch_format:
            return {"step": e.opt_step, "exp_avg": None if e.exp_avg is None else e.exp_avg.clone(),
                    "exp_avg_sq": None if e.exp_avg_sq is None else e.exp_avg_sq.clone(), "param_groups": groups}
        named, offset = self._named_trainable()
        state = {}
        if e.exp_avg is not None:
            for idx, name in named:
                state[idx] = {"step": torch.tensor(float(e.opt_step)),
                              "exp_avg": e.view(name, e.exp_avg).clone(), "exp_avg_sq": e.view(name, e.exp_avg_sq).clone()}
        g0 = dict(groups[0])
        g0.setdefault("amsgrad", False)
        g0["params"] = list(range(offset + len(named)))
        return {"state": state, "param_groups": [g0]}

    def load_state_dict(self, sd) -> None:
        e = self.engine
        if "state" in sd:                       # torch.optim.AdamW layout (ours or the reference's)
            named, _ = self._named_trainable()
            if sd["state"]:
                if e.exp_avg is None:
                    e.exp_avg = torch.zeros_like(e.params)
                    e.exp_avg_sq = torch.zeros_like(e.params)
                steps = set()
                for idx, name in named:
                    st = sd["state"].get(idx, sd["state"].get(str(idx)))
                    if st is None:
                        continue
                    e.view(name, e.exp_avg).copy_(st["exp_avg"].to(e.device))
                    e.view(name, e.exp_avg_sq).copy_(st["exp_avg_sq"].to(e.device))
                    steps.add(int(float(st["step"])))
                if len(steps) > 1:
                    raise ValueError("B200AdamW keeps one step counter; the checkpoint has per-parameter steps %s" % sorted(steps))
                e.opt_step = steps.pop() if steps else 0
                e._step_dev.fill_(e.opt_step)
            for g, s_ in zip(self.param_groups, sd["param_groups"]):
                g.update({k: v for k, v in s_.items() if k in ("lr", "betas", "eps", "weight_decay", "initial_lr")})
            return
        e.opt_step = int(sd["step"])
        e._step_dev.fill_(e.opt_step)
        if sd["exp_avg"] is not None:
            e.exp_avg = sd["exp_avg"].to(e.device).clone()
            e.exp_avg_sq = sd["exp_avg_sq"].to(e.device).clone()
        for g, s in zip(self.param_groups, sd["param_groups"]):
            g.update({k: v for k, v in s.items() if k != "params"})


# train_one_epoch packs the captions' non-PAD prefixes (var-len path of the engine) unless B200_PACKED=0
PACKED_BATCHES = os.environ.get("B200_PACKED", "1") != "0"


def trim_batch(decoder_input_tokens: torch.Tensor, target_tokens: torch.Tensor, pad_idx: int, multiple: int = 8):
    """Drop the all-PAD tail columns of a collated batch.  The reference pads every caption to
    MAX_SEQ_LEN (tokenizer.py:306, dataset.py:195-197), so most columns of a real batch are PAD for
    every sample; they contribute nothing to the loss (CrossEntropyLoss(ignore_index=PAD),
    train.py:327), are masked as attention keys (decoder.py:162) and their embedding-row gradient
    is zero (decoder.py:105), so cutting them changes neither loss nor gradients -- only the work.
    The kept length is rounded up to `multiple` to bound the number of distinct shapes."""
    T = decoder_input_tokens.shape[1]
    live = ((decoder_input_tokens != pad_idx) | (target_tokens != pad_idx)).any(dim=0)
    n = int(live.nonzero().max().item()) + 1 if bool(live.any()) else 1
    n = min(T, (n + multiple - 1) // multiple * multiple)
    return decoder_input_tokens[:, :n], target_tokens[:, :n]


def caption_lengths(tokens: torch.Tensor, pad_idx: int) -> torch.Tensor:
    """Number of columns up to and including the last non-PAD token of every row of a padded caption matrix
    (what tokenizer.py:293-313 produces: every caption padded / truncated to MAX_SEQ_LEN)."""
    live = tokens != pad_idx
    pos = torch.arange(1, tokens.shape[1] + 1, device=tokens.device)
    return (live * pos).amax(dim=1)


class LengthBucketBatchSampler:
    """`batch_sampler` for `torch.utils.data.DataLoader` (reference train.py:282-289 uses shuffle=True over captions
    that are all padded to MAX_SEQ_LEN): every epoch the indices are shuffled, cut into pools of `pool_batches`
    batches, sorted by caption length inside a pool and cut into batches, whose order is shuffled again.  The
    samples of a batch then have similar lengths, so `trim_batch` removes most of the PAD columns the random
    batches of the reference would keep (a batch is as wide as its longest caption) -- same samples per epoch, same
    loss definition, a fraction of the decoder work.  With `world_size` > 1 every rank takes the batches
    `rank, rank + world_size, ...` of the same seeded order (the data-parallel step reduces over ranks, so the
    ranks of a step should see similar widths: consecutive batches of a pool are given to consecutive ranks)."""

    def __init__(self, lengths, batch_size: int, pool_batches: int = 50, shuffle: bool = True, seed: int = 0,
                 drop_last: bool = False, rank: int = 0, world_size: int = 1):
        self.lengths = torch.as_tensor(lengths).to(torch.int64).cpu()
        assert self.lengths.dim() == 1 and batch_size >= 1 and pool_batches >= 1 and 0 <= rank < world_size
        self.batch_size, self.pool_batches, self.shuffle, self.seed = batch_size, pool_batches, shuffle, seed
        self.drop_last, self.rank, self.world_size = drop_last, rank, world_size
        self.epoch = 0

    def set_epoch(self, epoch: int) -> None:
        self.epoch = epoch

    def _batches(self):
        n = self.lengths.numel()
        g = torch.Generator().manual_seed(self.seed + self.epoch)
        order = torch.randperm(n, generator=g) if self.shuffle else torch.arange(n)
        pool = self.batch_size * self.pool_batches * self.world_size
        batches = []
        for p0 in range(0, n, pool):
            idx = order[p0:p0 + pool]
            idx = idx[torch.argsort(self.lengths[idx], stable=True)]
            chunk = [idx[i:i + self.batch_size] for i in range(0, idx.numel(), self.batch_size)]
            if chunk and chunk[-1].numel() < self.batch_size and self.drop_last:
                chunk.pop()
            # keep groups of world_size consecutive (similar-length) batches together, shuffle the groups
            groups = [chunk[i:i + self.world_size] for i in range(0, len(chunk), self.world_size)]
            if self.shuffle and len(groups) > 1:
                perm = torch.randperm(len(groups), generator=g).tolist()
                groups = [groups[i] for i in perm]
            for grp in groups:
                batches.extend(grp)
        usable = len(batches) - len(batches) % self.world_size     # every rank runs the same number of steps
        return [b.tolist() for b in batches[:usable][self.rank::self.world_size]] if self.world_size > 1 else [b.tolist() for b in batches]

    def __iter__(self):
        return iter(self._batches())

    def __len__(self) -> int:
        return len(self._batches())


def _ignore_index(criterion) -> int:
    return int(getattr(criterion, "ignore_index", config.PAD_TOKEN_ID))


_ZERO_STREAMS = {}


# the gradient-norm pass runs per bucket on a side stream under the rest of backward (B200_OVERLAP_NORM=0: one pass at the end)
OVERLAP_NORM = os.environ.get("B200_OVERLAP_NORM", "1") != "0"
_NORM_STATE = {}


def _norm_events(eng):
    """(one event per gradient bucket, side stream) of an engine; created OUTSIDE any graph capture.  torch creates the
    CUDA event lazily at the first record(), and the engine records these events through their raw handles, so every
    event is recorded once here (a wait on a never-created event would be a silent no-op)."""
    key = id(eng)
    if key not in _NORM_STATE:
        if torch.cuda.is_current_stream_capturing():
            return None
        evs = [torch.cuda.Event() for _ in eng.grad_buckets()]
        for ev in evs:
            ev.record()
        _NORM_STATE[key] = (evs, torch.cuda.Stream(device=eng.device))
    return _NORM_STATE[key]


def _zero_stream(device) -> "torch.cuda.Stream":
    key = (device.type, device.index)
    if key not in _ZERO_STREAMS:
        _ZERO_STREAMS[key] = torch.cuda.Stream(device=device)
    return _ZERO_STREAMS[key]


def fused_train_step(model, images, decoder_input_tokens, target_tokens, optimizer: B200AdamW, ignore_index: int,
                     grad_clip_value: float, dp: Optional[DataParallel] = None, lengths=None) -> torch.Tensor:
    """One optimisation step; returns the device tensor [loss, n_valid] (no host sync).
    lengths (host, B ints): packed / var-len step -- PAD positions are not computed at all (see decoder.loss)."""
    decoder = model.decoder if hasattr(model, "decoder") else model
    # the gradient arena is cleared on a side stream while the forward runs (the forward never touches it);
    # fork / join through stream events, so the pattern is also valid under CUDA-graph capture
    cur = torch.cuda.current_stream()
    side = _zero_stream(cur.device)
    side.wait_stream(cur)
    with torch.cuda.stream(side):
        optimizer.zero_grad()
    if hasattr(model, "decoder"):
        out = model.loss(images, decoder_input_tokens, target_tokens, ignore_index, training=True, lengths=lengths)
    else:   # a bare decoder: `images` is the memory
        out = decoder.loss(decoder_input_tokens, target_tokens, images, None, ignore_index, training=True, lengths=lengths)
    cur.wait_stream(side)
    eng = decoder.engine
    norm_ready = False
    if dp is not None and dp.world_size > 1:
        inv = dp.global_inv_count(out)                 # 1 / (non-PAD targets over all ranks)
        norm_ready = dp.backward_and_allreduce(inv)    # bucketed all-reduce (+ per-bucket norm) overlapped with backward
        out = dp.global_loss(out, inv)
    elif OVERLAP_NORM and eng.buckets_cover_arena() and _norm_events(eng) is not None:
        # the global gradient norm (clip_grad_norm_, train.py:97) bucket by bucket on a side stream, behind the event the
        # engine records when a bucket's gradients are final: the 288 MB pass leaves the critical path
        evs, nstream = _norm_events(eng)
        eng.norm_begin()
        decoder.backward(events=evs)
        with torch.cuda.stream(nstream):
            for i, ev in enumerate(evs):
                nstream.wait_event(ev)
                eng.norm_add_bucket(i)
        cur.wait_stream(nstream)
        norm_ready = True
    else:
        decoder.backward()
    optimizer.step(max_grad_norm=grad_clip_value, norm_ready=norm_ready)
    return out


class GraphedTrainStep:
    """Fused train step replayed from a CUDA graph (removes the ~2-4 us launch gap after
    each of the ~210 kernels of a step).  The first `warmup` calls run eagerly (they are real steps);
    the next call captures the step over static input buffers and replays it from then on.  The
    AdamW step counter and learning rate live on the device, so replays stay exact."""

    def __init__(self, model, optimizer: B200AdamW, ignore_index: int, grad_clip_value: float, warmup: int = 2,
                 dp: Optional[DataParallel] = None):
        self.model, self.optimizer = model, optimizer
        self.ignore_index, self.clip = ignore_index, grad_clip_value
        self.warmup, self.calls = warmup, 0
        self.dp = dp          # data-parallel: the bucketed NCCL all-reduces are captured as graph nodes too
        self.graph, self.static_in, self.static_out = None, None, None

    def _eager(self, images, tokens, targets):
        return fused_train_step(self.model, images, tokens, targets, self.optimizer, self.ignore_index, self.clip, self.dp)

    def __call__(self, images, tokens, targets) -> torch.Tensor:
        self.calls += 1
        if self.graph is None:
            if self.calls <= self.warmup:
                return self._eager(images, tokens, targets)
            # the tensors of the capturing call become the static inputs: later calls that pass the same
            # tensors (e.g. a staging slot refilled by an H2D copy) replay without any extra copy
            self.static_in = (images, tokens, targets)
            eng = (self.model.decoder if hasattr(self.model, "decoder") else self.model).engine
            lr = float(self.optimizer.param_groups[0]["lr"])
            if eng._lr_host != lr:
                eng._lr_dev.fill_(lr)
                eng._lr_host = lr
            _zero_stream(torch.cuda.current_stream().device)   # made outside the capture
            _norm_events(eng)
            torch.cuda.synchronize()
            self.graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.graph):
                self.static_out = self._eager(*self.static_in)
            # the graph holds raw pointers into the workspace and the dropout state: the engine refuses to
            # reallocate either from now on (a larger eval batch must be reserved BEFORE capture, engine.reserve_workspace)
            eng.pin_for_graph(self)
            self.graph.replay()
            return self.static_out
        for d, s in zip(self.static_in, (images, tokens, targets)):
            if d.data_ptr() != s.data_ptr():
                d.copy_(s, non_blocking=True)
        eng = (self.model.decoder if hasattr(self.model, "decoder") else self.model).engine
        lr = float(self.optimizer.param_groups[0]["lr"])     # a scheduler may have moved it since the last replay
        if eng._lr_host != lr:
            eng._lr_dev.fill_(lr)
            eng._lr_host = lr
        eng.opt_step += 1
        eng._shadow_fresh = True          # the replayed AdamW kernel rewrites the bf16 shadow
        self.graph.replay()
        return self.static_out

    def release(self) -> None:
        """Drop the captured graph (and the engine's allocation pin with it)."""
        eng = (self.model.decoder if hasattr(self.model, "decoder") else self.model).engine
        self.graph = None
        eng.unpin_for_graph(self)


def train_one_epoch(model, dataloader, optimizer, criterion, device, grad_clip_value, scheduler, epoch,
                    log_interval, wandb_run, dp: Optional[DataParallel] = None):
    """Mean training loss of one epoch (reference train.py:62-123)."""
    model.train()
    total_loss, num_batches = 0.0, len(dataloader)
    fused = isinstance(optimizer, B200AdamW)
    pad = int(getattr(model, "decoder_pad_idx", getattr(model, "pad_idx", config.PAD_TOKEN_ID)))
    for i, batch in enumerate(dataloader):
        images = batch["images"].to(device, non_blocking=True)
        tokens, targets = batch["decoder_input_tokens"], batch["target_tokens"]
        lengths = None
        if fused and _ignore_index(criterion) == pad:
            tokens, targets = trim_batch(tokens, targets, pad)      # on the host tensors: no device sync
            # var-len: the captions' non-PAD prefixes only (tokenizer.py:293-313 pads every caption to MAX_SEQ_LEN);
            # B200_PACKED=0 keeps the padded rectangle.  None when a caption has a PAD inside it.
            if PACKED_BATCHES and tokens.device.type == "cpu":
                from .engine import DecoderEngine
                lengths = DecoderEngine.packed_lengths(tokens, pad)
        tokens = tokens.to(device, non_blocking=True)
        targets = targets.to(device, non_blocking=True)
        if fused:
            out = fused_train_step(model, images, tokens, targets, optimizer, _ignore_index(criterion),
                                   grad_clip_value, dp, lengths=lengths)
            batch_loss = float(out[0].item())
        else:
            optimizer.zero_grad()
            logits = model(images, tokens)
            loss = criterion(logits.view(-1, logits.size(-1)), targets.reshape(-1))
            loss.backward()
            if grad_clip_value > 0:
                torch.nn.utils.clip_grad_norm_(model.parameters(), grad_clip_value)
            optimizer.step()
            batch_loss = loss.item()
        if scheduler:
            current_lr = scheduler.get_last_lr()[0]
            scheduler.step()
        else:
            current_lr = optimizer.param_groups[0]["lr"]
        total_loss += batch_loss
        if wandb_run and (epoch * num_batches + i + 1) % log_interval == 0:
            wandb_run.log({"train_batch_loss": batch_loss, "learning_rate": current_lr,
                           "global_step": epoch * num_batches + i + 1})
    return total_loss / max(num_batches, 1)


def evaluate(model, dataloader, criterion, device):
    """Mean evaluation loss (reference train.py:125-151); fused LM-head + CE, no logits."""
    model.eval()
    total_loss, num_batches = 0.0, len(dataloader)
    ii = _ignore_index(criterion)
    pad = int(getattr(model, "decoder_pad_idx", getattr(model, "pad_idx", config.PAD_TOKEN_ID)))
    with torch.no_grad():
        for batch in dataloader:
            images = batch["images"].to(device, non_blocking=True)
            tokens, targets = batch["decoder_input_tokens"], batch["target_tokens"]
            lengths = None
            if ii == pad:
                tokens, targets = trim_batch(tokens, targets, pad)
                if PACKED_BATCHES and tokens.device.type == "cpu":      # var-len path, as in train_one_epoch
                    from .engine import DecoderEngine
                    lengths = DecoderEngine.packed_lengths(tokens, pad)
            tokens = tokens.to(device, non_blocking=True)
            targets = targets.to(device, non_blocking=True)
            total_loss += float(model.loss(images, tokens, targets, ii, training=False, lengths=lengths)[0].item())
    return total_loss / max(num_batches, 1)
