"""Hand-computable micro cases that pin the [UPSTREAM] conventions of the oracle
(offset order, even-kernel origin, union, prune, dedup) — SURVEY.md §8c item 2."""
import pytest
import torch

import me_oracle as ME


def _st(coords, feats, **k):
    return ME.SparseTensor(torch.tensor(feats, dtype=torch.float32), torch.tensor(coords, dtype=torch.int32), **k)


def test_offset_enumeration_x_fastest_and_centred():
    o = ME.kernel_offsets(3, (1, 1, 1))
    assert o.shape == (27, 3)
    assert o[0].tolist() == [-1, -1, -1] and o[1].tolist() == [0, -1, -1] and o[3].tolist() == [-1, 0, -1]
    assert o[13].tolist() == [0, 0, 0] and o[26].tolist() == [1, 1, 1]
    e = ME.kernel_offsets(2, (2, 2, 2))
    assert e.tolist() == [[0, 0, 0], [2, 0, 0], [0, 2, 0], [2, 2, 0], [0, 0, 2], [2, 0, 2], [0, 2, 2], [2, 2, 2]]


def test_two_voxel_conv3_uses_the_right_weight_slices():
    # voxels a=(0,0,0), b=(1,0,0): out[a] = a·W[13] + b·W[14]; out[b] = b·W[13] + a·W[12]
    x = _st([[0, 0, 0, 0], [0, 1, 0, 0]], [[1.0], [10.0]])
    conv = ME.MinkowskiConvolution(1, 1, kernel_size=3, dimension=3)
    with torch.no_grad():
        conv.kernel.copy_(torch.arange(27, dtype=torch.float32).view(27, 1, 1))
    y = conv(x)
    assert y.F.flatten().tolist() == [1 * 13 + 10 * 14, 10 * 13 + 1 * 12]


def test_stride2_even_kernel_floors_negative_coordinates():
    # children of parent (-2,0,0): (-2,0,0)→k=0, (-1,0,0)→k=1 ; (-3,..) belongs to parent -4
    x = _st([[0, -2, 0, 0], [0, -1, 0, 0], [0, -3, 1, 1]], [[1.0], [2.0], [4.0]])
    conv = ME.MinkowskiConvolution(1, 1, kernel_size=2, stride=2, dimension=3)
    with torch.no_grad():
        conv.kernel.copy_(torch.tensor([1., 10., 100., 1000., 1e4, 1e5, 1e6, 1e7]).view(8, 1, 1))
    y = conv(x)
    got = {tuple(c.tolist()): f.item() for c, f in zip(y.C, y.F.flatten())}
    assert got == {(0, -2, 0, 0): 1 * 1 + 2 * 10, (0, -4, 0, 0): 4 * 1e7}
    assert y.tensor_stride == [2, 2, 2]


def test_generative_transpose_children_and_weights():
    x = _st([[0, 4, -2, 0]], [[3.0]], tensor_stride=2)
    up = ME.MinkowskiConvolutionTranspose(1, 1, kernel_size=2, stride=2, dimension=3, expand_coordinates=True)
    with torch.no_grad():
        up.kernel.copy_(torch.arange(1, 9, dtype=torch.float32).view(8, 1, 1))
    y = up(x)
    got = {tuple(c.tolist()): f.item() for c, f in zip(y.C, y.F.flatten())}
    exp = {}
    k = 0
    for dz in (0, 1):
        for dy in (0, 1):
            for dx in (0, 1):
                exp[(0, 4 + dx, -2 + dy, 0 + dz)] = 3.0 * (k + 1)
                k += 1
    assert got == exp


def test_union_add_and_same_key_add():
    a = _st([[0, 0, 0, 0], [0, 1, 0, 0]], [[1.0], [2.0]])
    b = ME.SparseTensor(torch.tensor([[10.0], [20.0]]), torch.tensor([[0, 1, 0, 0], [0, 5, 5, 5]], dtype=torch.int32),
                        coordinate_manager=a.coordinate_manager)
    u = a + b
    got = {tuple(c.tolist()): f.item() for c, f in zip(u.C, u.F.flatten())}
    assert got == {(0, 0, 0, 0): 1.0, (0, 1, 0, 0): 12.0, (0, 5, 5, 5): 20.0}
    s = a + a
    assert s.coordinate_map_key == a.coordinate_map_key and s.F.flatten().tolist() == [2.0, 4.0]


def test_prune_keeps_order_allows_empty_and_raises_on_mismatch():
    x = _st([[0, 3, 0, 0], [0, 1, 0, 0], [0, 2, 0, 0]], [[1.0], [2.0], [3.0]])
    pr = ME.MinkowskiPruning()
    y = pr(x, torch.tensor([True, False, True]))
    assert y.C[:, 1].tolist() == [3, 2] and y.F.flatten().tolist() == [1.0, 3.0]
    assert pr(x, torch.zeros(3, dtype=torch.bool)).F.shape == (0, 1)
    with pytest.raises(RuntimeError):
        pr(x, torch.ones(4, dtype=torch.bool))
    # a pruned tensor still convolves against its own (new) map
    conv = ME.MinkowskiConvolution(1, 1, kernel_size=3, dimension=3)
    assert conv(y).F.shape == (2, 1)


def test_duplicate_coordinates_first_wins():
    x = _st([[0, 1, 1, 1], [0, 2, 2, 2], [0, 1, 1, 1]], [[1.0], [2.0], [3.0]])
    assert x.F.flatten().tolist() == [1.0, 2.0] and x.C.shape == (2, 4)


def test_features_at_and_coordinates_at_drop_the_batch_column():
    x = _st([[0, 1, 1, 1], [1, 2, 2, 2], [0, 3, 3, 3]], [[1.0], [2.0], [3.0]])
    assert x.features_at(0).flatten().tolist() == [1.0, 3.0]
    assert x.coordinates_at(1).tolist() == [[2, 2, 2]]


def test_scatter_max_matches_loop_and_zero_fills_empty_segments():
    src = torch.tensor([[1., -5.], [3., -7.], [2., -1.]])
    idx = torch.tensor([2, 0, 2])
    out, arg = ME.scatter_max(src, idx, dim=0)
    assert out.tolist() == [[3., -7.], [0., 0.], [2., -1.]]
    assert arg[0].tolist() == [1, 1] and arg[2].tolist() == [2, 2]


def test_kernel1_conv_is_a_matmul_with_2d_weight():
    conv = ME.MinkowskiConvolution(3, 2, kernel_size=1, bias=True, dimension=3)
    assert conv.kernel.shape == (3, 2) and conv.bias.shape == (1, 2)
    x = _st([[0, 0, 0, 0]], [[1.0, 2.0, 3.0]])
    assert torch.allclose(conv(x).F, x.F @ conv.kernel + conv.bias)
