"""Host sampler mirror + oracle remaps vs arrays captured from the unmodified reference
(tests/golden/make_golden.py: reference sampler.py:90-160, preprocess.py:311-407)."""
import os

import numpy as np
import pytest

import oracle
from gnn_b200 import graphgen, sampler

CASES = ["cora_gcn", "tiny_sage3", "tiny_order0"]


def _load(golden_dir, name):
    return np.load(os.path.join(golden_dir, name + ".npz"))


@pytest.mark.parametrize("name", CASES)
def test_ladies_sample_matches_reference(golden_dir, name):
    z = _load(golden_dir, name)
    shape = graphgen.SHAPES[str(z["shape"])]
    g = graphgen.generate(shape, seed=0)
    orders = [int(o) for o in z["orders"]]
    world = int(z["world"])
    for si, seed in enumerate(z["seeds"]):
        pre = f"s{si}_"
        rank = int(z[pre + "rank"])
        mb = sampler.ladies_sample(int(seed), z[pre + "batch_nodes"], [int(z["samp_num"])] * 5, shape.num_nodes,
                                   g.indptr, g.indices, orders)
        assert len(mb.layers) == int(z[pre + "nlayers"])
        assert mb.input_nodes.size == int(z[pre + "n0"])
        for li, layer in enumerate(mb.layers):
            lp = pre + f"l{li}_"
            if layer is None:
                assert lp + "none" in z.files
                continue
            for k in ["fullrowptr", "rowptr", "colidx", "normfact"]:
                ref = z[lp + k]
                got = getattr(layer, k)
                assert got.dtype == ref.dtype, (k, got.dtype, ref.dtype)
                assert np.array_equal(got, ref), k
            assert layer.nrows == int(z[lp + "nrows"]) and layer.ncols == int(z[lp + "ncols"])
            assert np.array_equal(mb.sampled_nodes[li], z[lp + "sampled_nodes"])
        # placement remap, reference sampler.py:150-158
        did = z["device_id_of_nodes_group"][rank]
        idx = z["idx_of_nodes_on_device_group"][rank]
        devices = list(range(world))
        pr = sampler.placement_remap(mb.input_nodes, did, idx, devices)
        assert np.array_equal(pr.mask_on_cpu, z[pre + "mask_cpu"])
        assert np.array_equal(pr.idx_on_cpu, z[pre + "idx_cpu"])
        for i in range(world):
            assert np.array_equal(pr.mask_on_devices[i], z[pre + f"mask_dev{i}"])
            assert np.array_equal(pr.idx_on_devices[i], z[pre + f"idx_dev{i}"])
        # the compact (src_dev, slot) form is the same information
        o_src, o_slot = oracle.placement_remap(mb.input_nodes, did, idx, devices)
        assert np.array_equal(o_src, pr.src_dev) and np.array_equal(o_slot, pr.slot)
        for i in range(world):
            assert np.array_equal(o_src == i, z[pre + f"mask_dev{i}"])
            assert np.array_equal(o_slot[o_src == i], z[pre + f"idx_dev{i}"])
        assert np.array_equal(o_slot[o_src == -1], z[pre + "idx_cpu"])


@pytest.mark.parametrize("name", CASES)
def test_oracle_gather_matches_reference(golden_dir, name):
    z = _load(golden_dir, name)
    shape = graphgen.SHAPES[str(z["shape"])]
    feats = graphgen.features(shape, seed=1)
    world = int(z["world"])
    bufs = [feats[z["gpu_buffer_group"][i]] for i in range(world)]
    import hashlib
    for si in range(len(z["seeds"])):
        pre = f"s{si}_"
        n0 = int(z[pre + "n0"])
        src = np.full(n0, -1, np.int32)
        slot = np.zeros(n0, np.int64)
        slot[z[pre + "mask_cpu"]] = z[pre + "idx_cpu"]
        for i in range(world):
            src[z[pre + f"mask_dev{i}"]] = i
            slot[z[pre + f"mask_dev{i}"]] = z[pre + f"idx_dev{i}"]
        out = oracle.gather_rows(bufs, feats, src, slot)
        sha = np.frombuffer(hashlib.sha256(out.tobytes()).digest(), dtype=np.uint8)
        assert np.array_equal(sha, z[pre + "input_feat_sha"])
        if pre + "input_feat" in z.files:
            assert np.array_equal(out, z[pre + "input_feat"])


def test_oracle_sampled_nodes_direct():
    rng = np.random.Generator(np.random.PCG64(5))
    for _ in range(20):
        prev = rng.choice(5000, rng.integers(1, 300), replace=False)
        extra = rng.choice(5000, rng.integers(0, 800), replace=False)
        after = np.unique(np.concatenate([extra, prev]))
        assert np.array_equal(oracle.sampled_nodes(after, prev), np.where(np.isin(after, prev))[0])
        assert np.array_equal(sampler.sampled_nodes_remap(after, prev), np.where(np.isin(after, prev))[0])


def test_legacy_choice_on_support_equals_numpy_choice():
    """The device sampler evaluates numpy's legacy weighted draw on the support of p only; it must return exactly what
    RandomState.choice(N, size, p=p, replace=False) returns on the full-length p (reference sampler.py:128)."""
    from gnn_b200.gpu_sampler import legacy_choice_on_support, legacy_choice_on_support_numpy
    rng = np.random.Generator(np.random.PCG64(0))
    for trial in range(40):
        n = int(rng.integers(50, 20000))
        counts = rng.integers(0, 9, n) * (rng.random(n) < rng.uniform(0.05, 0.9))
        nz = np.flatnonzero(counts)
        if nz.size == 0:
            continue
        size = int(min(nz.size, rng.integers(1, n)))
        p = counts / counts.sum()
        a = np.random.RandomState(trial).choice(n, size, p=p, replace=False)
        rs = np.random.RandomState(trial)
        b = nz[legacy_choice_on_support(rs, p[nz], size)]
        assert np.array_equal(a, b), trial
        assert np.array_equal(a, nz[legacy_choice_on_support_numpy(np.random.RandomState(trial), p[nz], size)]), trial
        # and the generator state afterwards is the same (the next layer's draw continues the stream)
        rs2 = np.random.RandomState(trial)
        rs2.choice(n, size, p=p, replace=False)
        assert rs.random_sample() == rs2.random_sample()


def test_native_draw_on_sampler_sized_inputs_and_errors():
    """gnn_legacy_choice_f64 on LADIES-sized supports (heavy-tailed counts, several consecutive draws on one stream, like
    the layers of one minibatch) and its error return when p has fewer non-zero entries than the sample."""
    from gnn_b200.gpu_sampler import legacy_choice_on_support
    rng = np.random.Generator(np.random.PCG64(1))
    for n, size, seed in [(200000, 8192, 5), (30000, 8192, 7), (5000, 4999, 3), (10, 10, 1), (1000, 1, 123456789)]:
        counts = rng.zipf(1.6, n).clip(max=5000).astype(np.int64)
        p = counts / counts.sum()
        rs, ref = np.random.RandomState(seed), np.random.RandomState(seed)
        for _ in range(3):
            assert np.array_equal(legacy_choice_on_support(rs, p, size), ref.choice(n, size, p=p, replace=False))
    p = np.zeros(100)
    p[:5] = 0.2
    with pytest.raises(ValueError):
        legacy_choice_on_support(np.random.RandomState(0), p, 6)


def test_native_host_layer_equals_the_numpy_expressions():
    """gnn_ladies_layer_host (p, s_num, draw, union with previous_nodes, normfact, sampled_nodes remap in one native call)
    against the reference's numpy expressions on the same inputs, over several consecutive layers of one stream, with and
    without locality scaling (integer, dyadic and non-dyadic factors)."""
    from gnn_b200 import gpu_sampler as gs
    rng = np.random.Generator(np.random.PCG64(7))
    for trial, (n_nodes, n_nz, samp, scale) in enumerate([(50000, 30000, 8192, 1.0), (50000, 30000, 8192, 2.0), (200000, 9000, 8192, 1.5),
                                                         (3000, 50, 64, 16.0), (3000, 2000, 10000, 1.0), (100, 1, 5, 1.0),
                                                         # ids far sparser than the support: the sort/merge branch of the union
                                                         (3000000, 9000, 4096, 1.5), (5000000, 700, 8192, 1.0),
                                                         # the support is one contiguous id range (position = id - first id)
                                                         (30000, 30000, 8192, 1.0)]):
        rs = np.random.RandomState(100 + trial)
        state = gs.mt_state_of(np.random.RandomState(100 + trial))
        state_dense = state.copy()
        state_ex = state.copy()
        previous = rng.choice(n_nodes, size=min(512, n_nodes // 2), replace=False)
        for layer in range(3):
            nz = np.sort(rng.choice(n_nodes, size=n_nz, replace=False)).astype(np.int64)
            cnt = rng.zipf(1.7, n_nz).clip(max=4000).astype(np.int32)
            skew = np.unique(rng.choice(n_nodes, size=n_nodes // 10, replace=False)).astype(np.int64) if scale > 1 else None
            a = gs.host_layer_native(state, nz, cnt, skew, scale, previous, samp)
            b = gs.host_layer_numpy(rs, nz, cnt, skew, scale, previous, samp)
            assert np.array_equal(a[0], b[0]), (trial, layer, "after_nodes")
            assert np.array_equal(a[1].view(np.uint32), b[1].view(np.uint32)), (trial, layer, "normfact bits")
            assert np.array_equal(a[2], b[2]), (trial, layer, "sampled_nodes")
            assert a[3] == b[3]
            if n_nodes <= 200000:
                # the dense-count entry point (whole count array in, support compacted natively) gives the same answer
                dense = np.zeros(n_nodes, dtype=np.int32)
                dense[nz] = cnt
                c = gs.host_layer_native_dense(state_dense, dense, skew, scale, previous, samp)
                for u, v in zip(a[:3], c[:3]):
                    assert u.dtype == v.dtype and np.array_equal(u.view(np.uint8), v.view(np.uint8)), (trial, layer, "dense entry")
                assert a[3] == c[3]
                assert np.array_equal(state, state_dense)
                # and the entry that takes the compacted support AND the whole array (device-compacted supports)
                e = gs.host_layer_native(state_ex, nz, cnt, skew, scale, previous, samp, counts_dense=dense)
                for u, v in zip(a[:3], e[:3]):
                    assert u.dtype == v.dtype and np.array_equal(u.view(np.uint8), v.view(np.uint8)), (trial, layer, "ex entry")
                assert np.array_equal(state, state_ex)
            previous = a[0]
        # the generator ends in the same state
        assert np.array_equal(state, gs.mt_state_of(rs))


def test_native_host_layer_scaling_variants():
    """The three ways gnn_ladies_layer_host applies the locality scaling (sampler.py:119-121) - int32 values with a bitmap
    of the set, int64 values when count * scale would not fit, and the merge of two ascending lists for ids far sparser
    than the support - against the numpy expressions."""
    from gnn_b200 import gpu_sampler as gs
    rng = np.random.Generator(np.random.PCG64(11))
    for trial, (n_nodes, n_nz, n_skew_in, n_skew_out, scale) in enumerate([(60000, 20000, 3000, 3000, 3.0),           # bitmap, int32
                                                                            (60000, 20000, 3000, 3000, 2.0e6),         # bitmap, int64
                                                                            (400000000, 600, 80, 40, 2.5)]):           # merge
        rs = np.random.RandomState(300 + trial)
        state = gs.mt_state_of(np.random.RandomState(300 + trial))
        nz = np.sort(rng.choice(n_nodes, size=n_nz, replace=False)).astype(np.int64)
        cnt = rng.zipf(1.7, n_nz).clip(max=4000).astype(np.int32)
        skew = np.unique(np.concatenate((rng.choice(nz, size=n_skew_in, replace=False),
                                         rng.choice(n_nodes, size=n_skew_out, replace=False)))).astype(np.int64)
        previous = rng.choice(nz, size=min(300, n_nz // 2), replace=False)
        a = gs.host_layer_native(state, nz, cnt, skew, scale, previous, 256)
        b = gs.host_layer_numpy(rs, nz, cnt, skew, scale, previous, 256)
        assert np.array_equal(a[0], b[0]) and np.array_equal(a[2], b[2]) and a[3] == b[3], trial
        assert np.array_equal(a[1].view(np.uint32), b[1].view(np.uint32)), trial
        assert np.array_equal(state, gs.mt_state_of(rs)), trial


def test_native_host_layer_with_repeated_and_unsorted_previous_nodes():
    """A batch may repeat nodes and comes in any order (np.unique / np.in1d of sampler.py:131,143 do not care): dense and
    sparse id ranges, with and without the whole count array."""
    from gnn_b200 import gpu_sampler as gs
    rng = np.random.Generator(np.random.PCG64(21))
    for trial, (n_nodes, n_nz) in enumerate([(40000, 25000), (300000000, 900)]):
        nz = np.sort(rng.choice(n_nodes, size=n_nz, replace=False)).astype(np.int64)
        cnt = rng.zipf(1.8, n_nz).clip(max=3000).astype(np.int32)
        base = rng.choice(nz, size=200, replace=False)
        previous = np.concatenate((base, base[:50], rng.choice(n_nodes, size=20, replace=False)))     # repeats + ids off the support
        previous = previous[rng.permutation(previous.size)]
        rs = np.random.RandomState(500 + trial)
        state = gs.mt_state_of(np.random.RandomState(500 + trial))
        a = gs.host_layer_native(state, nz, cnt, None, 1.0, previous, 512)
        b = gs.host_layer_numpy(rs, nz, cnt, None, 1.0, previous, 512)
        assert np.array_equal(a[0], b[0]) and np.array_equal(a[2], b[2]) and a[3] == b[3], trial
        assert np.array_equal(a[1].view(np.uint32), b[1].view(np.uint32)), trial
        if n_nodes <= 100000:
            dense = np.zeros(n_nodes, dtype=np.int32)
            dense[nz] = cnt
            c = gs.host_layer_native_dense(gs.mt_state_of(np.random.RandomState(500 + trial)), dense, None, 1.0, previous, 512)
            assert all(np.array_equal(u.view(np.uint8), v.view(np.uint8)) for u, v in zip(a[:3], c[:3])), trial
