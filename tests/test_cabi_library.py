"""The C-ABI library loads without a GPU and exports exactly what include/rrt_b200.h declares.
No compute entry is called here (CPU box)."""
import ctypes as C
import os
import re

import pytest

from rrt_mil_b200 import cabi, build

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "rrt_b200.h")


@pytest.fixture(scope="module")
def lib():
    build.build()  # no-op when the in-tree .so is current
    return cabi.lib()


def declared_symbols():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"RRT_API\s+[\w\s\*]+?\b(rrt_\w+)\s*\(", text)))


def test_header_declares_what_the_binding_binds():
    assert declared_symbols() == sorted(cabi.SIGNATURES)


def test_every_declared_symbol_is_exported(lib):
    for name in declared_symbols():
        assert hasattr(lib, name), name


def test_abi_version_and_constants(lib):
    assert lib.rrt_abi_version() == cabi.RRT_ABI_VERSION
    text = open(HEADER).read()
    for macro, val in [("RRT_ABI_VERSION", cabi.RRT_ABI_VERSION),
                       ("RRT_MAX_RMSA_LAYERS", cabi.RRT_MAX_RMSA_LAYERS),
                       ("RRT_MAX_CRMSA_K", cabi.RRT_MAX_CRMSA_K),
                       ("RRT_MAX_EPEG_K", cabi.RRT_MAX_EPEG_K), ("RRT_MAX_LANES", cabi.RRT_MAX_LANES)]:
        assert int(re.search(rf"#define {macro} (\d+)", text).group(1)) == val


def test_grid_geometry_matches_oracle(lib):
    from oracle import rrt_oracle as O
    cases = [(L, g, rs, mn, mr) for L in (1, 2, 50, 63, 64, 65, 512, 576, 577, 9000, 9216, 50000, 123457)
             for (g, rs, mn, mr) in ((8, 0, 0, 0.0), (16, 0, 0, 0.0), (4, 0, 0, 0.0), (8, 5, 0, 0.0),
                                     (8, 0, 700, 0.0), (16, 0, 0, 5.0), (8, 0, 0, 0.5))]
    for L, g, rs, mn, mr in cases:
        H, r, _ = O.grid_geometry(L, g, rs, mn, mr)
        assert cabi.grid_geometry(L, g, rs, mn, mr) == (H, r), (L, g, rs, mn, mr)


def test_workspace_bytes_and_config_validation(lib):
    from rrt_mil_b200 import RRTEncoder
    cfg = RRTEncoder()._cfg
    n9000 = cabi.workspace_bytes(cfg, 9000)
    # fp16 z + qkv + o over 9216 padded tokens, two fp32 residual buffers over 9000: ~90 MB at D=512
    assert 70e6 < n9000 < 110e6
    assert cabi.workspace_bytes(cfg, 512) < n9000
    bad = cabi.RrtConfig.from_buffer_copy(cfg)
    bad.dim = 100
    with pytest.raises(ValueError):
        cabi.workspace_bytes(bad, 512)
    bad = cabi.RrtConfig.from_buffer_copy(cfg)
    bad.epeg_k = 4
    with pytest.raises(ValueError):
        cabi.workspace_bytes(bad, 512)
    with pytest.raises(ValueError):
        cabi.workspace_bytes(cfg, 0)
    assert b"geometry" in lib.rrt_last_error() or b"bag" in lib.rrt_last_error()


def test_struct_layout_matches_header_sizes(tmp_path):
    """The ctypes mirrors against the header itself: gcc compiles include/rrt_b200.h and prints the
    sizes and a few offsets of the POD structs."""
    import os, subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    src = tmp_path / "layout.c"
    src.write_text('''#include <stdio.h>
#include <stddef.h>
#include "rrt_b200.h"
int main(void) {
  printf("%zu %zu %zu %zu %zu %zu %zu %zu %zu\\n", sizeof(rrt_config), sizeof(rrt_attn_weights),
         sizeof(rrt_weights), sizeof(rrt_grads), offsetof(rrt_config, min_region_ratio),
         offsetof(rrt_config, pos), offsetof(rrt_weights, cr_attn), offsetof(rrt_weights, pos_w),
         offsetof(rrt_grads, cr_attn));
  return 0;
}
''')
    exe = tmp_path / "layout"
    subprocess.run(["gcc", "-I", os.path.join(root, "include"), str(src), "-o", str(exe)], check=True)
    got = [int(v) for v in subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split()]
    want = [C.sizeof(cabi.RrtConfig), C.sizeof(cabi.RrtAttnWeights), C.sizeof(cabi.RrtWeights),
            C.sizeof(cabi.RrtGrads), cabi.RrtConfig.min_region_ratio.offset, cabi.RrtConfig.pos.offset,
            cabi.RrtWeights.cr_attn.offset, cabi.RrtWeights.pos_w.offset, cabi.RrtGrads.cr_attn.offset]
    assert got == want


def test_integration_md_stub_structs_match_the_header():
    """The ctypes mirrors printed in INTEGRATION.md (what a reference maintainer would paste) must have the
    layout of the library's structs -- a stale field list there would corrupt memory silently."""
    import os, re
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    text = open(os.path.join(root, "INTEGRATION.md")).read()
    block = re.search(r"```python\n# modules/rrt_b200_stub.py.*?```", text, re.S).group(0)
    code = block[len("```python\n"):-3]
    code = code.split("def rrt_forward_b200")[0]                      # the struct definitions only
    code = code.replace('C.CDLL("librrt_b200.so")', "None").replace("import ctypes as C, torch", "import ctypes as C")
    ns = {}
    exec(code, ns)
    assert C.sizeof(ns["_Cfg"]) == C.sizeof(cabi.RrtConfig)
    assert C.sizeof(ns["_Attn"]) == C.sizeof(cabi.RrtAttnWeights)
    assert C.sizeof(ns["_Ffn"]) == C.sizeof(cabi.RrtFfnWeights)
    assert C.sizeof(ns["_W"]) == C.sizeof(cabi.RrtWeights)
    assert [n for n, _ in ns["_Cfg"]._fields_] == [n for n, *_ in cabi.RrtConfig._fields_]
    assert [n for n, _ in ns["_W"]._fields_] == [n for n, *_ in cabi.RrtWeights._fields_]


def test_config_validation_of_the_ablation_options():
    """pos / ffn options are validated by the library itself (no GPU needed: sizes only)."""
    from rrt_mil_b200 import RRTEncoder
    base = RRTEncoder()._cfg
    n0 = cabi.workspace_bytes(base, 9000)
    ffn = RRTEncoder(ffn=True)._cfg
    assert cabi.workspace_bytes(ffn, 9000) > n0 + 9000 * 2048 * 2          # the fp16 hidden rows
    peg = RRTEncoder(pos="ppeg", pos_pos=-1)._cfg
    assert cabi.workspace_bytes(peg, 9000) >= n0 + 9000 * 512 * 4           # the PEG output
    for field, bad in (("ffn_hidden", 100), ("ffn_act", 3), ("pos", 7), ("pos_pos", 2), ("peg_k", 4)):
        cfg = RRTEncoder(ffn=True, pos="peg", pos_pos=-1)._cfg
        setattr(cfg, field, bad)
        with pytest.raises(ValueError):
            cabi.workspace_bytes(cfg, 512)
    n = C.c_size_t()
    lib = cabi.lib()
    assert lib.rrt_train_tape_bytes(C.byref(base), 9000, C.byref(n)) == 0 and n.value > 9216 * 512 * 2 * 5
    assert lib.rrt_backward_workspace_bytes(C.byref(base), 9000, C.byref(n)) == 0 and n.value > 0
    assert lib.rrt_mil_head_backward_workspace_bytes(9000, 512, 128, C.byref(n)) == 0 and n.value > 9000 * 512 * 2
    assert lib.rrt_mil_head_backward_workspace_bytes(0, 512, 128, C.byref(n)) != 0


def test_optimizer_rejects_what_it_cannot_update():
    import torch
    from rrt_mil_b200.optim import Adam, AdamW
    with pytest.raises(ValueError):
        Adam([torch.nn.Parameter(torch.zeros(3))], lr=-1.0)
    with pytest.raises(ValueError):
        AdamW([torch.nn.Parameter(torch.zeros(3))], betas=(1.0, 0.9))
    p = torch.nn.Parameter(torch.zeros(3))
    p.grad = torch.ones(3)
    with pytest.raises(RuntimeError, match="CUDA"):      # no CPU fallback
        Adam([p]).step()
    q = torch.nn.Parameter(torch.zeros(3))
    Adam([q]).step()                                       # no gradient: nothing to do, no device needed
    assert float(q.abs().sum()) == 0.0
