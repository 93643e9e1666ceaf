"""CPU: the C-ABI library builds for sm_100a, loads, exports every symbol include/b2c.h declares, and the host
mirrors keep the reference's API surface.  No compute calls (there is no GPU here)."""
import ast
import ctypes
import inspect
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_SRC = "/root/reference/src"


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "b2c.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(b2c_[a-z0-9_]+)\s*\(", text)))


def test_header_and_binding_agree(lib):
    from imagecaptioner_b200 import _ops
    declared = _declared_symbols()
    assert declared, "no symbols parsed from include/b2c.h"
    assert sorted(_ops.SYMBOLS) == declared
    for name in declared:
        assert getattr(lib, name) is not None


def test_abi_version_and_error_string(lib):
    from imagecaptioner_b200 import _ops
    assert lib.b2c_abi_version() == _ops.ABI_VERSION == 5
    assert isinstance(lib.b2c_last_error(), bytes)


def test_workspace_bytes_is_pure_host_arithmetic(lib):
    from imagecaptioner_b200 import _ops
    s = _ops.B2CShape(512, 20, 49, 256, 512, 2, 5000)
    train = lib.b2c_workspace_bytes(ctypes.byref(s), _ops.B2C_BF16, _ops.B2C_WS_TRAIN)
    train32 = lib.b2c_workspace_bytes(ctypes.byref(s), _ops.B2C_F32, _ops.B2C_WS_TRAIN)
    dec = lib.b2c_workspace_bytes(ctypes.byref(s), _ops.B2C_BF16, _ops.B2C_WS_DECODE)
    assert 100e6 < train < 2e9 and train32 > train and 0 < dec < train
    bad = _ops.B2CShape(512, 20, 49, 250, 512, 2, 5000)          # E not a multiple of 8
    assert lib.b2c_workspace_bytes(ctypes.byref(bad), _ops.B2C_BF16, _ops.B2C_WS_TRAIN) == 0
    assert b"multiples of 8" in lib.b2c_last_error()


def test_sass_has_blackwell_instructions(lib):
    """tcgen05.mma -> UTCHMMA, tcgen05.ld -> LDTM, TMA -> UTMALDG / UBLKCP (B200_PROFILING.md)."""
    import shutil
    import subprocess
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        pytest.skip("cuobjdump not available")
    from imagecaptioner_b200 import _ops
    sass = subprocess.run([cuobjdump, "-sass", _ops.LIB_PATH], capture_output=True, text=True, check=True).stdout
    for mnemonic in ("UTCHMMA", "LDTM", "UTMALDG", "UBLKCP"):
        assert mnemonic in sass, mnemonic
    assert "sm_100a" in subprocess.run([cuobjdump, "-lelf", _ops.LIB_PATH], capture_output=True, text=True).stdout


def test_no_cpu_fallback():
    """The product path refuses CPU tensors instead of silently computing somewhere else."""
    from imagecaptioner_b200.student_model import LSTMDecoder
    from imagecaptioner_b200.distillation_utils import DistillationLoss
    dec = LSTMDecoder(50, 16, 32, 1, dropout=0.0)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        dec(torch.randn(2, 49, 16), torch.randint(0, 50, (3, 2)))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        DistillationLoss(vocab_size=50)({"logits": torch.randn(3, 2, 50)}, {"logits": torch.randn(3, 2, 50)}, torch.ones(3, 2, dtype=torch.long))


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "imagecaptioner_b200")
    for dirpath, _, files in os.walk(pkg):
        for fn in files:
            if fn.endswith(".py"):
                src = open(os.path.join(dirpath, fn)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), fn
                assert "/root/reference" not in src.replace("``/root/reference", ""), fn


def test_state_dict_keys_and_shapes_match_reference_layout():
    from imagecaptioner_b200.student_model import CaptioningStudent, PrecomputedFeatures
    m = CaptioningStudent(5000, 256, 512, 2, dropout=0.3, use_attention_refinement=True, encoder=PrecomputedFeatures())
    sd = m.state_dict()
    expect = {
        "decoder.embedding.weight": (5000, 256), "decoder.attention.weight": (256, 768), "decoder.attention.bias": (256,),
        "decoder.attention_combine.weight": (256, 512), "decoder.lstm.weight_ih_l0": (2048, 256), "decoder.lstm.weight_hh_l0": (2048, 512),
        "decoder.lstm.weight_ih_l1": (2048, 512), "decoder.lstm.bias_hh_l1": (2048,), "decoder.output_projection.0.weight": (256, 512),
        "decoder.output_projection.3.weight": (5000, 256), "decoder.output_projection.3.bias": (5000,),
        "attention_refinement.attention.in_proj_weight": (768, 256), "attention_refinement.ffn.3.weight": (256, 512),
        "attention_refinement.norm2.bias": (256,),
    }
    for k, shp in expect.items():
        assert tuple(sd[k].shape) == shp, k
    n_dec = sum(p.numel() for p in m.decoder.parameters())
    n_ref = sum(p.numel() for p in m.attention_refinement.parameters())
    assert n_dec == 6_702_728 and n_ref == 527_104          # SURVEY.md §8a (a1, a7), counted on the reference
    from oracle import kd_oracle as O
    theirs = O.init_student_params(5000, 256, 512, 2, True)
    assert set(theirs) == set(sd)


@pytest.mark.skipif(not os.path.isdir(REF_SRC), reason="reference tree not present (GPU box)")
@pytest.mark.parametrize("module", ["student_model", "distillation_utils"])
def test_public_surface_matches_reference(module):
    """Every class / function / method / argument name the reference module defines exists in the mirror."""
    import importlib
    tree = ast.parse(open(os.path.join(REF_SRC, module + ".py")).read())
    mine = importlib.import_module("imagecaptioner_b200." + module)
    for node in tree.body:
        if isinstance(node, ast.FunctionDef):
            fn = getattr(mine, node.name)
            assert [a.arg for a in node.args.args] == list(inspect.signature(fn).parameters)[: len(node.args.args)], node.name
        elif isinstance(node, ast.ClassDef):
            cls = getattr(mine, node.name)
            for item in node.body:
                if isinstance(item, ast.FunctionDef):
                    meth = getattr(cls, item.name)
                    ref_args = [a.arg for a in item.args.args]
                    my_args = list(inspect.signature(meth).parameters)
                    assert my_args[: len(ref_args)] == ref_args, f"{node.name}.{item.name}: {my_args} vs {ref_args}"
                    ref_defaults = [ast.literal_eval(d) for d in item.args.defaults]
                    my_defaults = [p.default for p in inspect.signature(meth).parameters.values() if p.default is not inspect._empty]
                    assert my_defaults[: len(ref_defaults)] == ref_defaults, f"{node.name}.{item.name} defaults"


def test_host_side_helpers():
    from imagecaptioner_b200.distillation_utils import compute_bleu_score, FeatureProjector

    class V:
        itos = {i: f"w{i}" for i in range(20)}
    assert compute_bleu_score([1, 5, 6, 7, 2, 0], [1, 5, 6, 9, 2], V) == pytest.approx(2 / 3)
    assert compute_bleu_score([5], [0, 1, 2], V) == 0.0
    fp = FeatureProjector(384, 256, 197, 64).eval()              # test_dimension_fix.py:16-43 (values: GPU test; shape pin: oracle test)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        fp(torch.randn(2, 197, 384))
    ident = FeatureProjector(384, 384, 197, 49)
    assert sum(p.numel() for p in ident.parameters()) == 0
