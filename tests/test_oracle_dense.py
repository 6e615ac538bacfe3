"""Pins the CPU oracle (oracle/me_oracle) against torch's own dense ops, which are exact
equivalents of the sparse ops on zero-filled grids.  Independent of
any code in this repository other than the oracle itself."""
import pytest
import torch
import torch.nn.functional as F

import me_oracle as ME

torch.manual_seed(0)


def _scene(shape=(8, 7, 5), p=0.4, C=6, batch=2, lo=(0, 0, 0), seed=0):
    g = torch.Generator().manual_seed(seed)
    coords, feats = [], []
    for b in range(batch):
        occ = torch.rand(*shape, generator=g) < p
        c = torch.nonzero(occ).int() + torch.tensor(lo, dtype=torch.int32)
        coords.append(c)
    bc = ME.utils.batched_coordinates(coords)
    f = torch.randn(bc.shape[0], C, generator=g, dtype=torch.float64)
    perm = torch.randperm(bc.shape[0], generator=g)      # row order must not matter
    return bc[perm], f[perm]


def _dense(x, shape, lo=(0, 0, 0)):
    B = int(x.C[:, 0].max()) + 1
    return x.dense(torch.Size([B, x.F.shape[1], *shape]), min_coordinate=torch.IntTensor(list(lo)))[0]


def _sample(dense, coords, lo=(0, 0, 0), stride=1):
    c = coords.long()
    return dense[c[:, 0], :, (c[:, 1] - lo[0]) // stride, (c[:, 2] - lo[1]) // stride, (c[:, 3] - lo[2]) // stride]


def _w_dense(kernel, k):
    """ME kernel [K,Cin,Cout] with x-fastest offsets → torch conv3d weight [Cout,Cin,kx,ky,kz]."""
    K, Cin, Cout = kernel.shape
    return kernel.view(k, k, k, Cin, Cout).permute(4, 3, 2, 1, 0).contiguous()   # [z,y,x]→[x,y,z]


@pytest.mark.parametrize("lo", [(0, 0, 0), (-8, 4, -2)])
def test_conv3_stride1_matches_conv3d(lo):
    shape = (8, 7, 5)
    C, F_ = _scene(shape, lo=lo)
    x = ME.SparseTensor(F_, C)
    conv = ME.MinkowskiConvolution(6, 4, kernel_size=3, dimension=3).double()
    y = conv(x)
    ref = F.conv3d(_dense(x, shape, lo), _w_dense(conv.kernel.detach(), 3), padding=1)
    assert torch.allclose(y.F.detach(), _sample(ref, y.C, lo), atol=1e-12)
    assert y.coordinate_map_key == x.coordinate_map_key


def test_conv3_backward_matches_conv3d():
    shape = (6, 6, 4)
    C, F_ = _scene(shape)
    F_.requires_grad_(True)
    x = ME.SparseTensor(F_, C)
    conv = ME.MinkowskiConvolution(6, 4, kernel_size=3, dimension=3).double()
    y = conv(x)
    g = torch.randn_like(y.F)
    (y.F * g).sum().backward()
    gF, gW = F_.grad.clone(), conv.kernel.grad.clone()

    Fd = F_.detach().clone().requires_grad_(True)
    xd = ME.SparseTensor(Fd, C)
    w = conv.kernel.detach().clone().requires_grad_(True)
    ref = F.conv3d(_dense(xd, shape), _w_dense(w, 3), padding=1)
    (_sample(ref, y.C) * g).sum().backward()
    assert torch.allclose(gF, Fd.grad, atol=1e-12)
    assert torch.allclose(gW, w.grad, atol=1e-12)


@pytest.mark.parametrize("lo", [(0, 0, 0), (-8, -4, -2)])
def test_conv2_stride2_matches_conv3d(lo):
    shape = (8, 6, 4)
    C, F_ = _scene(shape, lo=lo)
    x = ME.SparseTensor(F_, C)
    conv = ME.MinkowskiConvolution(6, 5, kernel_size=2, stride=2, dimension=3).double()
    y = conv(x)
    assert y.tensor_stride == [2, 2, 2]
    assert int((y.C[:, 1:] % 2).abs().sum()) == 0
    ref = F.conv3d(_dense(x, shape, lo), _w_dense(conv.kernel.detach(), 2), stride=2)
    assert torch.allclose(y.F.detach(), _sample(ref, y.C, lo, 2), atol=1e-12)
    # output coords = occupied 2³ blocks
    occ = F.max_pool3d((_dense(x, shape, lo).abs().sum(1, keepdim=True) > 0).double(), 2)
    assert y.F.shape[0] == int(occ.sum())


def test_generative_transpose_matches_conv_transpose3d():
    shape = (4, 3, 2)
    C, F_ = _scene(shape, p=0.6)
    C = C.clone()
    C[:, 1:] *= 2                                        # a stride-2 tensor
    x = ME.SparseTensor(F_, C, tensor_stride=2)
    up = ME.MinkowskiConvolutionTranspose(6, 3, kernel_size=2, stride=2, dimension=3,
                                          expand_coordinates=True).double()
    y = up(x)
    assert y.tensor_stride == [1, 1, 1]
    assert y.F.shape[0] == 8 * x.F.shape[0]
    xd = x.dense(torch.Size([2, 6, *shape]), min_coordinate=torch.IntTensor([0, 0, 0]))[0]
    w = up.kernel.detach().view(2, 2, 2, 6, 3).permute(3, 4, 2, 1, 0).contiguous()  # [Cin,Cout,x,y,z]
    ref = F.conv_transpose3d(xd, w, stride=2)
    assert torch.allclose(y.F.detach(), _sample(ref, y.C), atol=1e-12)


def test_maxpool_matches_max_pool3d():
    shape = (8, 8, 4)
    C, F_ = _scene(shape, p=0.5)
    F_ = F_.abs() + 0.1                                   # positive ⇒ zero-fill is neutral for max
    x = ME.SparseTensor(F_, C)
    for s in (2, 4):
        y = ME.MinkowskiMaxPooling(kernel_size=s, stride=s, dimension=3)(x)
        ref = F.max_pool3d(_dense(x, shape), s)
        assert torch.allclose(y.F, _sample(ref, y.C, stride=s))


def test_batchnorm_is_batchnorm1d_over_rows():
    C, F_ = _scene()
    x = ME.SparseTensor(F_.float(), C)
    bn = ME.MinkowskiBatchNorm(6)
    y = bn(x)
    ref = F.batch_norm(F_.float(), None, None, bn.bn.weight, bn.bn.bias, True, 0.1, 1e-5)
    assert torch.allclose(y.F, ref, atol=1e-6)
    sbn = ME.MinkowskiSyncBatchNorm.convert_sync_batchnorm(torch.nn.Sequential(bn))
    assert isinstance(sbn[0], ME.MinkowskiSyncBatchNorm) and isinstance(sbn[0].bn, torch.nn.SyncBatchNorm)
    assert list(sbn.state_dict().keys()) == list(torch.nn.Sequential(ME.MinkowskiBatchNorm(6)).state_dict().keys())


def test_dense_to_sparse_roundtrip_and_zero_drop():
    shape = (5, 4, 3)
    C, F_ = _scene(shape, batch=1, lo=(-4, 0, 2))
    F_[3] = 0.0                                           # an all-zero row disappears in to_sparse
    x = ME.SparseTensor(F_, C)
    d, mn, ts = x.dense(min_coordinate=torch.IntTensor([-4, 0, 2]))
    s = ME.to_sparse(d)
    assert s.F.shape[0] == x.F.shape[0] - 1
    keys = ME.pack_keys(s.C)
    assert torch.all(keys[1:] > keys[:-1])                # torch.where order == lexicographic (b,x,y,z)
    back = s.C.clone()
    back[:, 1:] += torch.tensor([-4, 0, 2], dtype=torch.int32)
    row = ME.lookup(x.C, back)
    assert torch.all(row >= 0) and torch.equal(x.F[row], s.F)
    with pytest.raises(ValueError):
        x.dense()                                         # negative coords need min_coordinate
