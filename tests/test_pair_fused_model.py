"""Index conventions of the fused real-FFT split passes of csrc/sparc_pair.cuh (numpy model, no GPU).

tools/model_pair_fused.py emulates, thread by thread, the pass structure of the two-frames-per-CTA transforms at
H = 2^11: the separate pre / post-processing passes and the fused forms (last forward radix-4 pass + split of the
mirror pairs, first transposed radix-4 pass built from the row block, plan 4,8,8,8).  Both must reproduce the dense
sub-sampled DCT block (reference sparc_public/sparc.py:648-701 sub_dct) and each other."""
import importlib.util
import os


def _model():
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools", "model_pair_fused.py")
    spec = importlib.util.spec_from_file_location("model_pair_fused", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def test_fused_passes_match_unfused_and_dense_dct():
    res = _model().main()
    assert abs(res["fwd_scale"] - 1.0) < 1e-9 and abs(res["tr_scale"] - 1.0) < 1e-9
    assert res["fwd_err_dense"] < 1e-9 and res["tr_err_dense"] < 1e-9          # pass structure = the DCT block
    assert res["fwd_err_fused"] < 1e-11 and res["tr_err_fused"] < 1e-11        # fused = unfused
    assert res["fwd_conflicts"] == 0                                             # fused forward pass: conflict free
    assert all(c == 0 for plan in res["pass_conflicts"] for c in plan)           # both swizzles, every pass
