"""CPU: the oracle restatement and the drop-in constructor against the FULL-WIDTH reference fixtures
(tests/golden/full_*.npz: F = 200, ffn 384, H*r = 96, MLP 512-256, L = 400, K = 60/100/148; oracle/make_golden_full.py)."""
import pytest
import torch

from oracle import ctr_oracle as O
from tests.helpers import FULL_CASES, full_batch, full_init, load_full, max_err_scaled, packed_err, packed_keys

TOL = 1e-4


@pytest.mark.parametrize("name", FULL_CASES)
def test_dropin_ctor_reproduces_reference_seeded_init(name):
    """The drop-in CTRModel built on the CPU under the fixture's seed has the reference's initial parameters (checksums);
    this is what lets the full-width fixtures omit the 30 MB of initial weights."""
    import b200ctr
    z, cfg = load_full(name)
    m = full_init(b200ctr.CTRModel, cfg, z)
    assert len(m.state_dict()) == len([k for k in z.files if k.startswith("initsum/")])


@pytest.mark.parametrize("name", ["full_k60_concat", "full_k148_s1"])
def test_oracle_matches_reference_at_full_width(name):
    import b200ctr
    z, cfg = load_full(name)
    g = cfg["golden"]
    m = full_init(b200ctr.CTRModel, cfg, z)
    sd = {k: v.detach().clone() for k, v in m.state_dict().items()}
    spec = O.ModelSpec(cfg, g["num_dim"], g["mask_dim"], cfg["data"]["cat_cols"])
    batch, y = full_batch(z, 0)
    loss, logits, _, aux, G, _ = O.loss_and_grads(sd, spec, batch, y)
    assert max_err_scaled(logits, torch.from_numpy(z["step0/logits"])) < TOL
    assert max_err_scaled(aux, torch.from_numpy(z["step0/aux"])) < TOL
    assert abs(float(loss) - float(z["step0/loss"])) < 1e-5
    for n in packed_keys(z, "grad0/"):
        e_sub, e_norm = packed_err(z, "grad0/" + n, G[n])
        assert e_sub < TOL and e_norm < TOL, (n, e_sub, e_norm)
