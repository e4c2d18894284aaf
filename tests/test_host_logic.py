"""CPU tests of the host side: C-ABI library loads and exports every symbol the header declares,
state_dict keys/shapes equal the reference's, integer subset logic, config errors, no silent fallback."""
import os
import re

import pytest
import torch

import multivae_b200 as mb
from multivae_b200 import _cabi
from multivae_b200.subsets import all_subsets, deterministic_selection, mvae_random_subsets
from oracle.cases import CASES

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")


def test_library_exports_every_declared_symbol():
    import __graft_entry__ as g
    g.build()
    hdr = open(os.path.join(ROOT, "include", "multivae_b200.h")).read()
    declared = set(re.findall(r"\b(mv_[a-z0-9_]+)\s*\(", hdr))
    lib = _cabi.lib()
    for sym in declared:
        assert hasattr(lib, sym), sym
    assert declared == set(_cabi.exported_symbols()), declared ^ set(_cabi.exported_symbols())
    mj, mn, sm = (_cabi.ctypes.c_int(), _cabi.ctypes.c_int(), _cabi.ctypes.c_int())
    assert lib.mv_version(mj, mn, sm) == 0 and sm.value == 100


def test_no_cpu_fallback():
    cfg = mb.MVTCAEConfig(n_modalities=2, latent_dim=4, input_dims={"a": (4,), "b": (6,)})
    model = mb.MVTCAE(cfg)
    ds = mb.MultimodalBaseDataset(data={"a": torch.rand(3, 4), "b": torch.rand(3, 6)})
    with pytest.raises(_cabi.NativeLibraryError):
        model(ds)


@pytest.mark.parametrize("name", sorted(CASES))
def test_state_dict_keys_match_reference(name):
    from tests.gpu_checks import MODELS, product_archs
    import copy
    rec = torch.load(os.path.join(GOLD, f"elbo_{name}.pt"), weights_only=False)
    spec = CASES[name]
    cls, cfgcls = MODELS[spec["model"]]
    model = cls(cfgcls(n_modalities=len(spec["dims"]), input_dims=dict(spec["dims"]), **copy.deepcopy(spec["cfg"])), *product_archs(spec))
    mine = {k: tuple(v.shape) for k, v in model.state_dict().items()}
    assert mine == rec["state_shapes"]
    ref_trainable = {k for k, g in rec["grads"].items() if g is not None or "prior" not in k}
    mine_trainable = {k for k, p in model.named_parameters() if p.requires_grad}
    assert {k for k in mine_trainable if "prior" in k} == {k for k in ref_trainable if "prior" in k}


@pytest.mark.parametrize("net", sorted(__import__("tests.net_checks", fromlist=["NETS"]).NETS))
def test_network_state_dict_matches_reference(net):
    from tests.net_checks import NETS
    rec = torch.load(os.path.join(GOLD, f"nets_{net}.pt"), weights_only=False)
    m = NETS[net]()
    assert {k: tuple(v.shape) for k, v in m.state_dict().items()} == rec["state_shapes"]


@pytest.mark.parametrize("net", sorted(__import__("tests.net_checks", fromlist=["NETS"]).NETS))
def test_network_forward_backward_matches_reference_on_cpu(net):
    """The product's nn modules (library-layer path, CPU) against outputs and gradients of the REAL reference modules."""
    from tests.net_checks import check_net
    check_net(net, "cpu", rtol=1e-4, atol=1e-5, grad_tol=1e-3)


def test_subset_logic_known_answers():
    assert list(all_subsets(["mod1", "mod2", "mod3"]).keys()) == ["", "mod1", "mod2", "mod3", "mod1_mod2", "mod1_mod3", "mod2_mod3", "mod1_mod2_mod3"]
    assert torch.bincount(deterministic_selection(256, 31).long()).tolist() == [8] * 30 + [16]
    assert torch.bincount(deterministic_selection(32, 31).long()).tolist() == [1] * 30 + [2]
    assert [list(s) for s in mvae_random_subsets(["a", "b", "c", "d"])] == [list(s) for s in __import__("itertools").combinations("abcd", 2)] + [list(s) for s in __import__("itertools").combinations("abcd", 3)]
    assert mvae_random_subsets(["a", "b"]) == []


def test_config_and_constructor_errors():
    with pytest.raises(AttributeError):
        mb.MMVAEPlus(mb.MMVAEPlusConfig(n_modalities=2, input_dims={"a": (4,), "b": (4,)}))  # modalities_specific_dim missing
    with pytest.raises(AttributeError):
        mb.MVTCAE(mb.MVTCAEConfig(n_modalities=3, input_dims={"a": (4,), "b": (4,)}))
    with pytest.raises(ValueError):
        mb.MMVAEConfig(n_modalities=2, loss="nope")
    cfg = mb.MVAEConfig(n_modalities=2, input_dims={"a": (4,), "b": (4,)}, k=3)
    assert mb.MVAE(cfg).k == 0  # k forced to 0 when M <= 2 (mvae_model.py:40-41)


def test_ctypes_structs_match_the_c_header(tmp_path):
    """The ctypes mirrors in _cabi.py must have the size and field offsets of the structs in include/multivae_b200.h
    (compiled here with the host C compiler)."""
    import ctypes
    import shutil
    import subprocess
    from multivae_b200 import _cabi
    cc = shutil.which("gcc") or shutil.which("cc")
    if cc is None:
        pytest.skip("no host C compiler")
    structs = {"mv_tapgemm_args": _cabi.TapGemmArgs, "mv_pack_item": _cabi.PackItem, "mv_unpack_item": _cabi.UnpackItem}
    lines = ['#include <stdio.h>', '#include <stddef.h>', f'#include "{os.path.join(ROOT, "include", "multivae_b200.h")}"', "int main(void) {"]
    for cname, cls in structs.items():
        lines.append(f'  printf("{cname} %zu\\n", sizeof({cname}));')
        for fname, _ in cls._fields_:
            lines.append(f'  printf("{cname}.{fname} %zu\\n", offsetof({cname}, {fname}));')
    lines += ["  return 0;", "}"]
    src = tmp_path / "layout.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "layout"
    subprocess.run([cc, str(src), "-o", str(exe)], check=True)
    got = dict(l.split() for l in subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.splitlines())
    for cname, cls in structs.items():
        assert int(got[cname]) == ctypes.sizeof(cls), cname
        for fname, _ in cls._fields_:
            assert int(got[f"{cname}.{fname}"]) == getattr(cls, fname).offset, (cname, fname)


def test_likelihoods_evaluator_batches_the_test_set_in_order():
    """metrics/likelihoods/likelihoods.py:13-61: the evaluator walks the test set in order in batches of eval_config.batch_size and
    divides the summed estimate by the number of datapoints (the estimator itself is stubbed here: no GPU)."""
    import multivae_b200 as mb

    class Stub(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.seen = []

        def compute_joint_nll(self, inputs, K, batch_size_K):
            self.seen.append((inputs.data["a"][:, 0].tolist(), K, batch_size_K))
            return inputs.data["a"].sum()

    ds = mb.MultimodalBaseDataset(data={"a": torch.arange(10.0).view(10, 1), "b": torch.zeros(10, 2)})
    m = Stub()
    ev = mb.LikelihoodsEvaluator(m, ds, eval_config=mb.LikelihoodsEvaluatorConfig(batch_size=4, num_samples=12, batch_size_k=6))
    ev.device = "cpu"
    out = ev.eval()
    assert [s[0] for s in m.seen] == [[0.0, 1.0, 2.0, 3.0], [4.0, 5.0, 6.0, 7.0], [8.0, 9.0]]
    assert all(s[1:] == (12, 6) for s in m.seen)
    assert float(out.joint_likelihood) == 4.5
    assert ev.joint_nll_from_subset(["a"]) is None
