"""CPU checks of the oracle's dropout restatement (oracle.DropSpec): torch dropout semantics
(mask then 1/(1-p) scaling at the reference's sites), determinism, site / counter independence."""
import torch

from oracle import decoder_oracle as O
from tests.helpers import CFGS, synth


def test_mask_statistics_and_independence():
    spec = O.DropSpec(0.1, 42, 1)
    idx = torch.arange(1 << 18, dtype=torch.int64)
    k0, k1 = spec.keep(0, idx), spec.keep(1, idx)
    assert abs(1 - k0.float().mean().item() - 0.1) < 4e-3
    assert abs(1 - k1.float().mean().item() - 0.1) < 4e-3
    both = (~k0 & ~k1).float().mean().item()            # independent sites: P(both dropped) ~ p^2
    assert abs(both - 0.01) < 2e-3
    k0b = O.DropSpec(0.1, 42, 2).keep(0, idx)           # next step: new masks
    assert (k0 != k0b).float().mean().item() > 0.1
    assert torch.equal(k0, O.DropSpec(0.1, 42, 1).keep(0, idx))
    x = torch.ones(8, 64)
    y = spec.rows(5, x)
    assert set(y.unique().tolist()) <= {0.0, 1.0 / 0.9} or torch.allclose(y[y > 0], torch.tensor(1 / 0.9))


def test_dropout_changes_loss_but_p0_is_identity():
    c = CFGS["nano"]
    p = O.init_params(c["V"], c["E"], c["H"], c["L"], c["F"], c["ML"], seed=1)
    tok, tgt, mem, _ = synth(c, 2)
    with torch.no_grad():
        base = O.decoder_forward(p, tok, mem, None, c["H"])
        same = O.decoder_forward(p, tok, mem, None, c["H"], drop=O.DropSpec(0.0, 3, 1))
        diff = O.decoder_forward(p, tok, mem, None, c["H"], drop=O.DropSpec(0.1, 3, 1))
    assert torch.equal(base, same)
    assert (base - diff).abs().max().item() > 1e-3
    l, g = O.loss_and_grads(p, tok, tgt, mem, None, c["H"], drop=O.DropSpec(0.1, 3, 1))
    assert torch.isfinite(l) and all(torch.isfinite(v).all() for v in g.values())
