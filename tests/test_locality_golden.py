"""--locality_sampling (BASELINE configs[3]) against vectors captured from the UNMODIFIED reference
(tests/golden/make_locality_golden.py): the skew node sets of preprocess.get_skewed_sampled_nodes and the LADIES
hand-off arrays of sampler.ladies_sampler with scale_factor 2, 1.5 and 16 (sampler.py:119-121)."""
import os

import numpy as np
import pytest

from gnn_b200 import graphgen, placement, sampler


@pytest.fixture(scope="module")
def z(golden_dir):
    return np.load(os.path.join(golden_dir, "locality_small.npz"))


def _setup(z, tag):
    shape = graphgen.SHAPES[str(z[tag + "_shape"])]
    g = graphgen.generate(shape, seed=0)
    world = int(z[tag + "_world"])
    orders = [int(o) for o in z[tag + "_orders"]]
    # the fixture's lap_matrix came from the reference's row_normalize on float32 ones (GraphSAGE) or on adj + eye,
    # which scipy promotes to float64 (GCN); the placement ranks nodes by sums of these values, so the dtype matters
    pl = placement.create_placement(g.to_scipy(np.float64 if shape.self_loops else np.float32), g.train_nodes, int(z[tag + "_buffer_size"]), list(range(world)),
                                    sum(orders), alpha=0.0)
    return shape, g, orders, pl


@pytest.mark.parametrize("tag", ["gcn", "sage"])
def test_skew_sets_match_reference(z, tag):
    shape, g, orders, pl = _setup(z, tag)
    assert np.array_equal(np.stack(pl.gpu_buffer_group), z[tag + "_gpu_buffer_group"])
    sets = placement.locality_sampling_sets(g.indptr, g.indices, shape.self_loops, pl.gpu_buffer_group, len(orders))
    for i, s in enumerate(sets):
        assert np.array_equal(s, z[f"{tag}_set{i}"]), (tag, i)


@pytest.mark.parametrize("tag", ["gcn", "sage"])
def test_scaled_sampler_matches_reference(z, tag):
    shape, g, orders, pl = _setup(z, tag)
    sets = [z[f"{tag}_set{i}"] for i in range(len(orders))]
    samp = int(z[tag + "_samp_num"])
    for ci in range(3):
        c = f"{tag}_c{ci}_"
        mb = sampler.ladies_sample(int(z[c + "seed"]), z[c + "batch_nodes"], [samp] * 5, shape.num_nodes, g.indptr, g.indices, orders,
                                   skewed_sampling_nodes=sets, scale_factor=float(z[c + "scale_factor"]))
        assert mb.input_nodes.size == int(z[c + "n0"])
        for li, layer in enumerate(mb.layers):
            for k in ["fullrowptr", "rowptr", "colidx"]:
                assert np.array_equal(getattr(layer, k), z[c + f"l{li}_{k}"]), (tag, ci, li, k)
            assert np.array_equal(layer.normfact.view(np.uint32), z[c + f"l{li}_normfact"].view(np.uint32)), (tag, ci, li)
            assert np.array_equal(mb.sampled_nodes[li], z[c + f"l{li}_sampled_nodes"])


def test_scale_factor_controller_follows_the_reference_branches():
    """placement.ScaleFactorController against a literal transcription of the (disabled) loop of reference main.py:200-212,
    driven by the same sequences of (data movement, execution) times."""
    from gnn_b200.placement import ScaleFactorController

    def reference_trace(ratios, scale_factor=1.0):
        factor_increase, factor_before, factor_after, out = True, scale_factor, scale_factor, []
        for r in ratios:
            if factor_increase == True:          # noqa: E712  (the reference's spelling)
                if scale_factor >= 16:
                    factor_increase = False
                elif r >= 0.2:
                    factor_before = scale_factor
                    scale_factor *= 2
                elif r < 0.1 and scale_factor != 1:
                    factor_after = scale_factor
                    scale_factor = (factor_before + factor_after) / 2
                else:
                    factor_increase = False
            out.append(scale_factor)
        return out

    rng = np.random.Generator(np.random.PCG64(3))
    cases = [[0.5, 0.4, 0.3, 0.05, 0.05, 0.3], [0.5] * 8, [0.05, 0.5], [0.15, 0.5], [0.3, 0.15, 0.5], [0.25, 0.25, 0.05, 0.05, 0.05]]
    cases += [list(rng.uniform(0.0, 0.4, 10)) for _ in range(20)]
    for ratios in cases:
        c = ScaleFactorController(1.0)
        got = [c.update(r * 7.0, 7.0) for r in ratios]
        assert got == reference_trace(ratios), ratios
