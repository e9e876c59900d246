"""CPU tests of the skeleton tables against fixtures produced by the reference's own kinematic classes
(tests/golden/make_reachability.py -> tests/golden/skeletons.npz): node names, node types, adjacency and the reachability
matrices of `covariance_matrix_type='reachability'` (kinematic/base.py:85-127), including AMASS-MANO (51 nodes, 43 types)."""
import os

import numpy as np
import pytest
import torch

import skeletondiffusion_b200 as sdb

Z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "skeletons.npz"))


@pytest.mark.parametrize("name", ["amass", "amass-mano", "h36m", "freeman"])
def test_skeleton_tables_match_reference(name):
    spec, key = sdb.get_skeleton(name), name.replace("-", "_")
    assert list(spec.node_names) == [str(s) for s in Z[f"{key}__names"]]
    assert list(spec.node_types) == Z[f"{key}__types"].tolist()
    assert np.array_equal(spec.adj_matrix.numpy(), Z[f"{key}__adj"])


@pytest.mark.parametrize("name", ["amass", "h36m", "freeman", "amass-mano"])
def test_reachability_matrix_is_bit_exact(name):
    spec, key = sdb.get_skeleton(name), name.replace("-", "_")
    stops = ("hips", "bmn", None) if name != "amass-mano" else ("hips",)       # the 51-node search takes ~10 s per variant
    for stop in stops:
        assert np.array_equal(spec.reachability_matrix(factor=0.5, stop_at=stop).numpy(), Z[f"{key}__reach_{stop}"]), stop
    if name != "amass-mano":
        assert np.array_equal(spec.reachability_matrix(factor=0.3, stop_at="hips").numpy(), Z[f"{key}__reach_hips_f03"])
    with pytest.raises(AssertionError):
        spec.reachability_matrix(stop_at=0)              # the reference's DiffusionManager default asserts as well


def test_diffusion_manager_with_reachability_covariance():
    """DiffusionManager(covariance_matrix_type='reachability') (diffusion_manager.py:20-22) builds a valid nonisotropic process."""
    spec = sdb.get_skeleton("h36m")
    mgr = sdb.DiffusionManager(diffusion_type="NonisotropicGaussianDiffusion", skeleton=spec, covariance_matrix_type="reachability",
                               reachability_matrix_degree_factor=0.5, reachability_matrix_stop_at="hips", num_nodes=spec.num_nodes,
                               node_types=spec.nodes_type_id, diffusion_conditioning=True, latent_size=96, diffusion_timesteps=10,
                               diffusion_arch=dict(depth=1, attn_heads=4, attn_dim_head=32, learn_influence=True))
    diff = mgr.get_diffusion()
    assert torch.allclose(diff.U @ torch.diag(diff.Lambda_N) @ diff.U_transposed, diff.Sigma_N, atol=1e-5)
    lam = diff.Lambda_N
    assert float(lam.max()) == pytest.approx(1.0, abs=1e-6) and float(lam.min()) > 0
    assert tuple(diff.posterior_mean_coef1_x0.shape) == (10, spec.num_nodes, spec.num_nodes)


def test_amass_mano_models_have_the_reference_shapes():
    spec = sdb.get_skeleton("amass-mano")
    ae, diff = sdb.build_models(spec, "cpu", depth=1)
    sd = diff.state_dict()
    assert tuple(sd["model.init_lin.weight"].shape) == (43, 192, 192) and tuple(sd["U"].shape) == (51, 51)
    assert tuple(ae.state_dict()["decoder.rnn.layers.0.weight_hh"].shape) == (43, 288, 96)
