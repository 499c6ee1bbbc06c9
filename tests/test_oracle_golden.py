"""Pins the CPU oracle (oracle/rrt_oracle.py) against the golden vectors generated from the
unmodified reference, and -- in the build container, where /root/reference exists -- against
the live reference module.  CPU only."""
import pytest
import torch

from oracle import rrt_oracle as O
from oracle import _reference_shim as shim
from golden_util import CASES, load_case, assert_matches_golden

# the 50k-token case needs ~2 GB in float64 reference order; keep it but only in "spec" order
BIG = {"c4_n50000_g16"}


@pytest.mark.parametrize("name", sorted(CASES))
@pytest.mark.parametrize("order", ["reference", "spec"])
def test_oracle_matches_golden(name, order):
    if name in BIG and order == "reference":
        pytest.skip("reference-order restatement of the 50k bag is covered by the spec order")
    cfg, w, x, gold = load_case(name)
    y = O.encoder_forward(x, w, cfg, order)
    # fixtures store float32 rows/row checksums: 1e-6 is storage rounding, not oracle slack
    assert_matches_golden(y, gold, 1e-6, f"oracle[{order}] {name}")


@pytest.mark.parametrize("name", ["c1_n512_d512", "d256_g4", "tiny_n50"])
def test_oracle_float32_close_to_float64(name):
    cfg, w, x, gold = load_case(name, dtype=torch.float32)
    y = O.encoder_forward(x, w, cfg, "spec")
    assert_matches_golden(y, gold, 2e-5, f"oracle fp32 {name}")


@pytest.mark.skipif(not shim.available(), reason="/root/reference only exists in the build container")
@pytest.mark.parametrize("L,over", [
    (333, dict()),
    (2048, dict(region_num=16, epeg_k=21, crmsa_k=5)),
    (100, dict(mlp_dim=128, n_heads=4, crmsa_heads=1, crmsa_mlp=True, all_shortcut=True)),
])
def test_oracle_matches_live_reference(L, over):
    cfg = O.EncoderConfig(**over)
    w = O.make_weights(cfg, 11)
    x = O.make_bag(L, cfg.mlp_dim, 12)
    m = shim.build_reference_encoder(cfg, w)
    with torch.no_grad():
        ref = m(x[None])[0]
    for order in ("reference", "spec"):
        assert O.rel_err(O.encoder_forward(x, w, cfg, order), ref) < 1e-12


def test_grid_geometry_known_values():
    # config sizes
    assert O.grid_geometry(9000, 8) == (96, 12, 216)
    assert O.grid_geometry(512, 8) == (24, 3, 64)
    assert O.grid_geometry(50000, 16) == (224, 14, 176)
    assert O.grid_geometry(50000, 8) == (224, 28, 176)
    assert O.grid_geometry(1, 8) == (8, 1, 63)
    assert O.grid_geometry(64, 8) == (8, 1, 0)
    assert O.grid_geometry(65, 8) == (16, 2, 191)


def test_region_slot_map_is_permutation_and_matches_view_permute():
    H, rs = 12, 3
    m = O.region_slot_map(H, rs)
    assert sorted(m.tolist()) == list(range(H * H))
    g = H // rs
    t = torch.arange(H * H).view(1, g, rs, g, rs).permute(0, 1, 3, 2, 4).reshape(-1)
    assert torch.equal(m, t)
