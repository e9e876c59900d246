"""CPU: the C-ABI library loads and exports every symbol include/skeldiff_b200.h declares (no compute
calls without a GPU); host-side logic (state_dict compatibility, window sharding, plan keys)."""
import os
import re
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "skeldiff_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(sd_[a-z0-9_]+)\s*\(", text)))


@pytest.fixture(scope="module")
def built_lib():
    import __graft_entry__ as ge
    ge.build()
    from skeletondiffusion_b200 import _native as nv
    return nv


def test_header_symbols_are_exported_and_bound(built_lib):
    nv = built_lib
    declared = _declared_symbols()
    assert len(declared) >= 25
    assert sorted(nv.SIGNATURES) == declared, "ctypes table and header disagree"
    lib = nv.load()
    for name in declared:
        assert hasattr(lib, name), f"{name} not exported"
    out = subprocess.run(["nm", "-D", "--defined-only", str(nv.library_path())], capture_output=True, text=True).stdout
    exported = set(re.findall(r"\bT (sd_[a-z0-9_]+)", out))
    assert set(declared) <= exported
    assert lib.sd_version() >= 100


def test_library_is_sm100a_only(built_lib):
    out = subprocess.run(["cuobjdump", "-lelf", str(built_lib.library_path())], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_(\d+a?)", out))
    assert archs == {"100a"}, archs


def test_missing_cuda_is_loud(built_lib):
    """No CPU fallback: CPU tensors are rejected before any native call."""
    import skeletondiffusion_b200 as sdb
    nv = built_lib
    model = sdb.Denoiser(dim=8, cond_dim=0, out_dim=8, channels=4, num_nodes=4, attn_heads=2, attn_dim_head=16)
    with pytest.raises(nv.NativeError):
        model(torch.zeros(2, 4, 8), torch.zeros(2, dtype=torch.long))
    corr = torch.eye(4)
    s, l, u = sdb.get_cov_from_corr(corr, if_run_as_isotropic=True)
    diff = sdb.NonisotropicGaussianDiffusion(Sigma_N=s, Lambda_N=l, U=u, model=model, latent_size=8)
    with pytest.raises(nv.NativeError):
        diff.sample(batch_size=2)


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "skeletondiffusion_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in src.replace("oracle's", ""), f"{f} mentions the oracle"


def test_state_dict_keys_match_survey_listing():
    import skeletondiffusion_b200 as sdb
    spec = sdb.get_skeleton("amass")
    ae, diff = sdb.build_models(spec, "cpu")
    sd = diff.state_dict()
    assert sum(p.numel() for p in diff.model.parameters()) == 32_135_979
    assert sum(p.numel() for p in ae.parameters()) == 2_100_207
    for k in ("betas", "alphas_cumprod", "Lambda_N", "Sigma_N", "U", "U_transposed", "posterior_mean_coef1_x0",
              "posterior_mean_coef2_xt", "Lambda_posterior_log_variance_clipped", "mahalanobis_S_sqrt_recip", "loss_weight",
              "inv_sqrt_Lambda_bar_mmUt", "inv_sqrt_Lambda_bar_sqrt_alphas_cumprod_mmUt", "Umm_sqrt_Lambda_bar_t",
              "Umm_sqrt_Lambda_bar_t_sqrt_recip_alphas_cumprod", "Lambda_posterior",
              "model.init_lin.G", "model.time_mlp.1.weight", "model.layers.0.0.mlp.1.weight", "model.layers.0.0.block1.proj.weight",
              "model.layers.0.1.fn.norm.g", "model.layers.6.1.fn.fn.to_qkv.weight", "model.layers.6.1.fn.fn.to_out.G",
              "model.final_res_block.res_linear.weight", "model.final_glin.bias"):
        assert k in sd, k
    assert "model.layers.7.1.fn.norm.g" not in sd                              # last pair has nn.Identity
    assert tuple(sd["model.layers.3.1.fn.fn.to_qkv.weight"].shape) == (13, 768, 192)
    assert tuple(sd["model.final_res_block.block1.proj.weight"].shape) == (13, 192, 384)
    asd = ae.state_dict()
    assert tuple(asd["decoder.rnn.layers.0.weight_ih"].shape) == (13, 288, 99)
    assert "decoder.rnn.layers.0.G_add" in asd and "encoder.rnn.layers.1.node_type_index" in asd
    assert "encoder.rnn.layers.0.G_add" not in asd


def test_cov_from_corr_matches_golden_and_isotropic_degenerate():
    import skeletondiffusion_b200 as sdb
    from tests import _golden as G
    case = G.load_npz("readme_perturbed")
    tabs = G.tables_of(case)
    s, l, u = sdb.get_cov_from_corr(case["corr"])
    assert torch.allclose(s, tabs["Sigma_N"], atol=5e-6) and torch.allclose(l, tabs["Lambda_N"], atol=5e-6)
    assert torch.allclose((u * l) @ u.T, s, atol=1e-5)
    assert abs(float(l.max()) - 1.0) < 1e-6
    s, l, u = sdb.get_cov_from_corr(case["corr"], if_run_as_isotropic=True)
    assert torch.equal(u, torch.eye(16)) and torch.equal(l, torch.ones(16)) and float(s.abs().max()) == 0.0


def test_diffusion_buffers_match_reference_tables():
    from tests import _golden as G
    for name in ("amass_perturbed", "amass_perturbed_iso", "h36m_perturbed"):
        case = G.load_npz(name)
        tabs = G.tables_of(case)
        import skeletondiffusion_b200 as sdb
        spec = sdb.get_skeleton(str(case["dataset"]))
        _, diff = sdb.build_models(spec, "cpu", if_run_as_isotropic=bool(case["iso"]))
        sd = diff.state_dict()
        # Eigenvectors are defined up to sign and LAPACK's choice depends on the host CPU: bring the fixture's U to this
        # host's signs (column j of U, hence row j of every "...mmUt" table and column j of every "Umm..." table), then
        # compare.  fp32 eigh itself differs by a few ulp between hosts (1.1e-6 seen on Lambda_N = 2.4e-5 relative
        # at the smallest eigenvalue 0.046, which the inverse-square-root tables inherit): 2e-5 of the table's scale.
        sgn = torch.sign((sd["U"] * tabs["U"]).sum(0))
        assert float(sgn.abs().min()) == 1.0
        for k, v in tabs.items():
            assert tuple(sd[k].shape) == tuple(v.shape), k
            if k == "U" or k.startswith("Umm"):
                v = v * sgn
            elif k in ("U_transposed", "mahalanobis_S_sqrt_recip") or k.endswith("mmUt"):   # S_sqrt_recip = diag(.) U^T
                v = v * sgn[:, None]
            assert float((sd[k] - v).abs().max()) <= 2e-5 * max(1.0, float(v.abs().max())), k


def test_shard_windows_partitions_exactly():
    import skeletondiffusion_b200 as sdb
    for n in (0, 1, 7, 512, 4096, 4099):
        for ws in (1, 2, 4, 8):
            spans = [sdb.shard_windows(n, r, ws) for r in range(ws)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1


def _gloo_worker(rank, world, port, ret):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import skeletondiffusion_b200 as sdb
    from skeletondiffusion_b200.distributed import gather_window_metrics
    n = 11
    lo, hi = sdb.shard_windows(n, rank, world)
    local = {"ade": torch.arange(lo, hi, dtype=torch.float32) * 2.0, "apd": torch.arange(lo, hi, dtype=torch.float32) + 0.5}
    full = gather_window_metrics(local, n, rank, world)
    ok = torch.equal(full["ade"], torch.arange(n, dtype=torch.float32) * 2.0) and torch.equal(full["apd"], torch.arange(n, dtype=torch.float32) + 0.5)
    ret[rank] = bool(ok)
    dist.destroy_process_group()


def test_metric_gather_world_size_2_gloo():
    import torch.multiprocessing as mp
    mgr = mp.Manager()
    ret = mgr.dict()
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_gloo_worker, args=(2, port, ret), nprocs=2, join=True)
    assert ret[0] and ret[1]


def test_reference_checkpoint_files_load_unchanged(tmp_path):
    """Checkpoint files written by the REFERENCE's own modules (torch.save({'model': state_dict, ...}), src/core/trainer.py /
    src/utils/load.py:11-17) load into the drop-in modules with strict keys, and files written from OUR modules load into the
    reference modules (oracle/_ref: present in the build container; skipped on a box without it)."""
    import torch
    from oracle import make_ref
    if not make_ref.add_to_path():
        pytest.skip("oracle/_ref absent")
    import contextlib, io
    import skeletondiffusion_b200 as sdb
    import bench
    spec = sdb.get_skeleton("h36m")
    with contextlib.redirect_stdout(io.StringIO()):
        ref_ae, ref_diff, _ = bench._reference_models(spec, "h36m", False, torch.device("cpu"))
    ae_path, diff_path = str(tmp_path / "checkpoint_100_val.pt"), str(tmp_path / "checkpoint_300_val.pt")
    torch.save({"model": ref_ae.state_dict(), "epoch": 100, "optimizer": {}}, ae_path)
    torch.save({"model": ref_diff.state_dict(), "epoch": 300, "ema": None}, diff_path)
    (ae, diff), dev = sdb.prepare_model(spec, ae_path, diff_path, device="cpu")
    for k, v in ref_diff.state_dict().items():
        assert torch.equal(diff.state_dict()[k], v), k
    for k, v in ref_ae.state_dict().items():
        assert torch.equal(ae.state_dict()[k], v), k
    # the other direction: a checkpoint written from our modules loads into the reference's, strict
    torch.save({"model": diff.state_dict()}, diff_path)
    ref_diff.load_state_dict(sdb.load_model_checkpoint(diff_path)["model"], strict=True)


def test_gate_interleaved_gru_rows_reproduce_the_cell():
    """Host logic of the fused GRU-step kernels (sd_gru_set_fused / sd_gru_set_fused_f16x2): with the rows of weight_ih / weight_hh
    and the biases permuted by gate_interleave_perm, every 96-column block of the products holds gates r | z | n of 32 hidden units,
    and gating block by block gives the reference cell's new state (oracle gru_cell, recurrent.py:333-358).  CPU, float64."""
    from oracle import skeldiff_oracle as oc
    from skeletondiffusion_b200.plan import gate_interleave_perm
    H, IN, N, B = 96, 7, 5, 4
    g = torch.Generator().manual_seed(11)
    sd = {"weight_ih": torch.randn(3 * H, IN, generator=g, dtype=torch.float64) * 0.3, "weight_hh": torch.randn(3 * H, H, generator=g, dtype=torch.float64) * 0.1,
          "bias_ih": torch.randn(3 * H, generator=g, dtype=torch.float64) * 0.1, "bias_hh": torch.randn(3 * H, generator=g, dtype=torch.float64) * 0.1}
    x = torch.randn(B, N, IN, generator=g, dtype=torch.float64)
    h = torch.tanh(torch.randn(B, N, H, generator=g, dtype=torch.float64))
    ref = oc.gru_cell(sd, "", x, h, torch.eye(N, dtype=torch.float64), None)
    perm = gate_interleave_perm(H)
    assert sorted(perm.tolist()) == list(range(3 * H))
    xr = x @ sd["weight_ih"][perm].t() + sd["bias_ih"][perm]
    hr = h @ sd["weight_hh"][perm].t() + sd["bias_hh"][perm]
    out = torch.empty_like(h)
    for blk in range(H // 32):
        xg, hg = xr[..., 96 * blk:96 * blk + 96], hr[..., 96 * blk:96 * blk + 96]
        r = torch.sigmoid(xg[..., :32] + hg[..., :32])
        z = torch.sigmoid(xg[..., 32:64] + hg[..., 32:64])
        n = torch.tanh(xg[..., 64:] + r * hg[..., 64:])
        out[..., 32 * blk:32 * blk + 32] = n - n * z + z * h[..., 32 * blk:32 * blk + 32]
    assert torch.allclose(out, ref, atol=1e-12, rtol=0)
    with pytest.raises(ValueError):
        gate_interleave_perm(100)
