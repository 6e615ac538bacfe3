"""CPU ORACLE — test infrastructure, NOT product code.

A plain-PyTorch (CPU, fp32/fp64) restatement of the MinkowskiEngine v0.5.4
semantics that PaSCo's hot path relies on.  MinkowskiEngine is a
third-party dependency that is absent from /root/reference (pinned only in prose,
/root/reference/README.md:90), so this file restates its published algorithm:

  * coordinate map  = dedup of int32 (b,x,y,z) rows (first occurrence wins)
  * kernel map      = for every kernel offset k, pairs (in_row, out_row) with
                      in_coord == out_coord + offset_k          (cross-correlation)
  * convolution     = for k: out[out_k] += in[in_k] @ W[k]     (gather → GEMM → scatter-add)

and anchors parity on the reference's own call sites:
  pasco/maskpls/mink.py:505-534,618-658   (conv / generative transpose / residual block)
  pasco/models/decoder_v3.py:148-172      (prune, concat coords, union add)
  pasco/models/unet3d_sparse_v2.py:182-214 (dense(), to_sparse, re-insert in the same manager)
  pasco/models/augmenter.py:13-27         (dense(min_coordinate) / to_sparse)
  pasco/models/transformer/transformer_predictor_v2.py:220-289 (max pooling, dense)
  pasco/loss/criterion_sparse.py:273-297  (features_at / coordinates_at)

PARITY UNPINNED: the reference ships no tests, golden vectors or fixtures for this
path and MinkowskiEngine itself cannot be built or imported here.
What pins this oracle instead is dense equivalence against torch's own
F.conv3d / F.conv_transpose3d / F.max_pool3d / F.batch_norm (tests/test_oracle_dense.py)
and hand-computed micro cases (tests/test_oracle_micro.py).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this package.  The product (pasco_b200/) never does.
"""
from __future__ import annotations

import itertools
from typing import List, Optional, Sequence, Tuple, Union

import numpy as np
import torch
import torch.nn as nn

__version__ = "0.5.4-oracle"

# ----------------------------------------------------------------------------
# integer core: packed keys, unique, lookups
# ----------------------------------------------------------------------------
_BIAS = 1 << 15  # coordinates may be negative (SURVEY.md §7 "hard parts")


def pack_keys(coords: torch.Tensor) -> torch.Tensor:
    """(b,x,y,z) int rows → one int64 key, order-preserving for lexicographic sort."""
    c = coords.to(torch.int64)
    return (((c[:, 0] + _BIAS) << 48) | ((c[:, 1] + _BIAS) << 32)
            | ((c[:, 2] + _BIAS) << 16) | (c[:, 3] + _BIAS))


def unique_first(coords: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """Dedup rows keeping the first occurrence and the input order.

    Returns (unique_index [N'], inverse [N]) like ME's insert_and_map."""
    n = coords.shape[0]
    if n == 0:
        z = torch.zeros(0, dtype=torch.int64)
        return z, z
    keys = pack_keys(coords)
    sk, order = torch.sort(keys, stable=True)
    first = torch.ones(n, dtype=torch.bool)
    first[1:] = sk[1:] != sk[:-1]
    group = torch.cumsum(first.to(torch.int64), 0) - 1          # group id in sorted order
    first_pos = order[first]                                    # first occurrence of each group
    # rank groups by first occurrence so the output keeps input order
    rank_order = torch.argsort(first_pos, stable=True)
    unique_index = first_pos[rank_order]
    group_rank = torch.empty_like(rank_order)
    group_rank[rank_order] = torch.arange(rank_order.numel())
    inverse = torch.empty(n, dtype=torch.int64)
    inverse[order] = group_rank[group]
    return unique_index, inverse


def lookup(table_coords: torch.Tensor, query_coords: torch.Tensor) -> torch.Tensor:
    """Row of each query coordinate in table_coords, −1 when absent."""
    nt = table_coords.shape[0]
    if nt == 0 or query_coords.shape[0] == 0:
        return torch.full((query_coords.shape[0],), -1, dtype=torch.int64)
    tk = pack_keys(table_coords)
    stk, order = torch.sort(tk)
    qk = pack_keys(query_coords)
    pos = torch.searchsorted(stk, qk).clamp_(max=nt - 1)
    hit = stk[pos] == qk
    return torch.where(hit, order[pos], torch.full_like(pos, -1))


def kernel_offsets(kernel_size: int, tensor_stride: Sequence[int], dilation: int = 1) -> torch.Tensor:
    """[K,3] offsets; axis 0 (x) fastest; odd k centred, even k in [0,k)  [UPSTREAM]."""
    if kernel_size % 2 == 1:
        r = list(range(-(kernel_size // 2), kernel_size // 2 + 1))
    else:
        r = list(range(kernel_size))
    offs = []
    for z, y, x in itertools.product(r, r, r):   # x fastest
        offs.append((x * tensor_stride[0] * dilation, y * tensor_stride[1] * dilation,
                     z * tensor_stride[2] * dilation))
    return torch.tensor(offs, dtype=torch.int64)


def neighbour_table(in_coords: torch.Tensor, out_coords: torch.Tensor, offsets: torch.Tensor) -> torch.Tensor:
    """nbr[k, o] = row i of in_coords with in_coords[i] == out_coords[o] + offsets[k], else −1."""
    K = offsets.shape[0]
    no = out_coords.shape[0]
    nbr = torch.full((K, no), -1, dtype=torch.int64)
    oc = out_coords.to(torch.int64)
    for k in range(K):
        q = oc.clone()
        q[:, 1:] += offsets[k]
        nbr[k] = lookup(in_coords, q)
    return nbr


def floor_to_stride(coords: torch.Tensor, stride: Sequence[int]) -> torch.Tensor:
    c = coords.clone().to(torch.int64)
    s = torch.tensor(list(stride), dtype=torch.int64)
    c[:, 1:] = torch.div(c[:, 1:], s, rounding_mode="floor") * s
    return c.to(torch.int32)


# ----------------------------------------------------------------------------
# coordinate manager
# ----------------------------------------------------------------------------
class CoordinateMapKey:
    _counter = itertools.count()

    def __init__(self, tensor_stride: Sequence[int], name: str = ""):
        self.tensor_stride = tuple(int(s) for s in tensor_stride)
        self.name = name
        self.uid = next(CoordinateMapKey._counter)

    def get_tensor_stride(self):
        return list(self.tensor_stride)

    def get_key(self):
        return (list(self.tensor_stride), self.name)

    def get_coordinate_size(self):
        return 4

    def __hash__(self):
        return hash(self.uid)

    def __eq__(self, other):
        return isinstance(other, CoordinateMapKey) and other.uid == self.uid

    def __repr__(self):
        return f"coordinate map key:{list(self.tensor_stride)}:{self.name or self.uid}"


class CoordinateManager:
    def __init__(self, D: int = 3, **_):
        self.D = D
        self._coords = {}
        self._stride_cache = {}
        self._kmap_cache = {}

    # -- maps ---------------------------------------------------------------
    def insert_and_map(self, coords: torch.Tensor, tensor_stride=1, string_id: str = ""):
        ts = _triple(tensor_stride)
        coords = coords.detach().cpu().to(torch.int32)
        uidx, inv = unique_first(coords)
        key = CoordinateMapKey(ts, string_id)
        self._coords[key] = coords[uidx].contiguous()
        return key, (uidx, inv)

    def _register(self, coords: torch.Tensor, ts) -> CoordinateMapKey:
        key = CoordinateMapKey(ts)
        self._coords[key] = coords.to(torch.int32).contiguous()
        return key

    def get_coordinates(self, key: CoordinateMapKey) -> torch.Tensor:
        return self._coords[key]

    def size(self, key) -> int:
        return self._coords[key].shape[0]

    def number_of_unique_batch_indices(self) -> int:
        allb = torch.cat([c[:, 0] for c in self._coords.values()])
        return int(torch.unique(allb).numel())

    def stride(self, in_key: CoordinateMapKey, stride) -> CoordinateMapKey:
        s = _triple(stride)
        ck = (in_key, s)
        if ck not in self._stride_cache:
            new_ts = tuple(a * b for a, b in zip(in_key.tensor_stride, s))
            c = floor_to_stride(self._coords[in_key], new_ts)
            uidx, _ = unique_first(c)
            self._stride_cache[ck] = self._register(c[uidx], new_ts)
        return self._stride_cache[ck]

    def generate(self, in_key: CoordinateMapKey, kernel_size: int, stride) -> CoordinateMapKey:
        """Generative transposed conv output: every input emits kernel_size³ children
        at in + off·out_stride (reference use: mink.py:524-527, k=2, s=2)."""
        s = _triple(stride)
        out_ts = tuple(a // b for a, b in zip(in_key.tensor_stride, s))
        offs = kernel_offsets(kernel_size, out_ts)
        c = self._coords[in_key].to(torch.int64)
        allc = []
        for k in range(offs.shape[0]):
            q = c.clone()
            q[:, 1:] += offs[k]
            allc.append(q)
        allc = torch.cat(allc, 0).to(torch.int32)
        uidx, _ = unique_first(allc)
        return self._register(allc[uidx], out_ts)

    def kernel_map(self, in_key, out_key, kernel_size: int, dilation: int = 1, transpose: bool = False):
        """nbr[K, N_out] table of input rows (−1 = no neighbour)."""
        ck = (in_key, out_key, kernel_size, dilation, transpose)
        if ck not in self._kmap_cache:
            cin, cout = self._coords[in_key], self._coords[out_key]
            if not transpose:
                offs = kernel_offsets(kernel_size, in_key.tensor_stride, dilation)
                nbr = neighbour_table(cin, cout, offs)
            else:
                # out (fine) row o receives in (coarse) row i through W[k] iff
                # out = in + off_k·out_stride   ⇔   in = out − off_k
                offs = kernel_offsets(kernel_size, out_key.tensor_stride, dilation)
                nbr = neighbour_table(cin, cout, -offs)
            self._kmap_cache[ck] = nbr
        return self._kmap_cache[ck]

    def union_map(self, keys: Sequence[CoordinateMapKey]):
        cs = [self._coords[k] for k in keys]
        allc = torch.cat(cs, 0)
        uidx, inv = unique_first(allc)
        out_key = self._register(allc[uidx], keys[0].tensor_stride)
        maps, start = [], 0
        for c in cs:
            n = c.shape[0]
            maps.append((torch.arange(n), inv[start:start + n]))
            start += n
        return out_key, maps


def _triple(v) -> Tuple[int, int, int]:
    if isinstance(v, (list, tuple)):
        assert len(v) == 3
        return tuple(int(x) for x in v)
    if isinstance(v, torch.Tensor):
        v = v.flatten().tolist()
        if len(v) == 1:
            v = v * 3
        return tuple(int(x) for x in v)
    return (int(v),) * 3


# ----------------------------------------------------------------------------
# SparseTensor
# ----------------------------------------------------------------------------
class SparseTensorQuantizationMode:
    RANDOM_SUBSAMPLE = 0
    UNWEIGHTED_AVERAGE = 1
    UNWEIGHTED_SUM = 2
    NO_QUANTIZATION = 3
    MAX_POOL = 4


class MinkowskiAlgorithm:
    DEFAULT = 0
    MEMORY_EFFICIENT = 1
    SPEED_OPTIMIZED = 2


class TensorField:  # only referenced in isinstance checks (pasco/models/dropout.py:23,47)
    pass


class SparseTensor:
    def __init__(self, features: torch.Tensor, coordinates: Optional[torch.Tensor] = None,
                 tensor_stride=1, coordinate_map_key: Optional[CoordinateMapKey] = None,
                 coordinate_manager: Optional[CoordinateManager] = None,
                 quantization_mode=SparseTensorQuantizationMode.RANDOM_SUBSAMPLE,
                 allocator_type=None, minkowski_algorithm=None, requires_grad=None, device=None):
        assert isinstance(features, torch.Tensor) and features.ndim == 2, "features must be [N, C]"
        if coordinate_map_key is None:
            assert coordinates is not None, "coordinates or coordinate_map_key required"
            assert coordinates.ndim == 2 and coordinates.shape[1] == 4, "coordinates must be [N, 4]"
            assert coordinates.shape[0] == features.shape[0], "coordinates / features length mismatch"
            assert not coordinates.dtype.is_floating_point, "coordinates must be integer"
            if coordinate_manager is None:
                coordinate_manager = CoordinateManager()
            coordinate_map_key, (uidx, inv) = coordinate_manager.insert_and_map(coordinates, tensor_stride)
            if uidx.numel() != features.shape[0]:
                features = features[uidx.to(features.device)]
        else:
            assert coordinate_manager is not None
            assert coordinate_manager.size(coordinate_map_key) == features.shape[0], \
                "features do not match the coordinate map size"
        self._F = features
        self.coordinate_manager = coordinate_manager
        self.coordinate_map_key = coordinate_map_key
        if requires_grad is not None:
            self._F.requires_grad_(requires_grad)

    # -- attributes -----------------------------------------------------------
    @property
    def F(self):
        return self._F

    @property
    def features(self):
        return self._F

    @property
    def C(self):
        return self.coordinate_manager.get_coordinates(self.coordinate_map_key).to(self._F.device)

    @property
    def coordinates(self):
        return self.C

    @property
    def tensor_stride(self):
        return list(self.coordinate_map_key.tensor_stride)

    @property
    def D(self):
        return 3

    @property
    def dimension(self):
        return 3

    @property
    def shape(self):
        return self._F.shape

    def size(self, *a):
        return self._F.size(*a)

    @property
    def device(self):
        return self._F.device

    @property
    def dtype(self):
        return self._F.dtype

    @property
    def requires_grad(self):
        return self._F.requires_grad

    def __len__(self):
        return self._F.shape[0]

    def __repr__(self):
        return f"SparseTensor(oracle, F={tuple(self._F.shape)}, stride={self.tensor_stride})"

    # -- per-batch accessors (criterion_sparse.py:273-274) -----------------------
    def _batch_rows(self, b: int):
        return torch.nonzero(self.C[:, 0] == b, as_tuple=True)[0]

    def features_at(self, b: int):
        return self._F[self._batch_rows(b)]

    def coordinates_at(self, b: int):
        return self.C[self._batch_rows(b)][:, 1:]

    @property
    def decomposed_features(self):
        nb = int(self.C[:, 0].max()) + 1 if len(self) else 0
        return [self.features_at(b) for b in range(nb)]

    @property
    def decomposed_coordinates(self):
        nb = int(self.C[:, 0].max()) + 1 if len(self) else 0
        return [self.coordinates_at(b) for b in range(nb)]

    # -- dense ----------------------------------------------------------------
    def dense(self, shape=None, min_coordinate=None, contract_stride=True):
        C = self.C
        ts = torch.tensor(self.tensor_stride, dtype=torch.int32, device=C.device)
        batch = C[:, 0].long()
        if min_coordinate is None:
            min_coordinate = C.min(0, keepdim=True)[0][:, 1:]
            if not torch.all(min_coordinate >= 0):
                raise ValueError(f"Coordinate has a negative value: {min_coordinate}. "
                                 "Please provide min_coordinate argument")
            coords = C[:, 1:]
        elif isinstance(min_coordinate, int) and min_coordinate == 0:
            coords = C[:, 1:]
        else:
            assert isinstance(min_coordinate, torch.Tensor) and not min_coordinate.dtype.is_floating_point
            min_coordinate = min_coordinate.to(C.device)
            if min_coordinate.ndim == 1:
                min_coordinate = min_coordinate.unsqueeze(0)
            coords = C[:, 1:] - min_coordinate
        assert int((torch.as_tensor(min_coordinate).to(ts.device) % ts).sum()) == 0, \
            "The minimum coordinates must be divisible by the tensor stride."
        if contract_stride:
            coords = torch.div(coords, ts, rounding_mode="floor")
        nch = self._F.shape[1]
        if shape is None:
            size = coords.max(0)[0] + 1
            shape = torch.Size([int(batch.max()) + 1, nch, *[int(s) for s in size]])
        else:
            assert len(shape) == 5 and shape[1] == nch
        out = torch.zeros(tuple(int(s) for s in shape), dtype=self._F.dtype, device=self._F.device)
        t = coords.long()
        out[batch, :, t[:, 0], t[:, 1], t[:, 2]] = self._F
        return out, min_coordinate, torch.tensor(self.tensor_stride, dtype=torch.int32)

    # -- arithmetic -------------------------------------------------------------
    def _binary(self, other, fn):
        if isinstance(other, (int, float)) or (isinstance(other, torch.Tensor)):
            return SparseTensor(fn(self._F, other), coordinate_map_key=self.coordinate_map_key,
                                coordinate_manager=self.coordinate_manager)
        assert isinstance(other, SparseTensor)
        assert other.coordinate_manager is self.coordinate_manager, "different coordinate managers"
        if other.coordinate_map_key == self.coordinate_map_key:
            return SparseTensor(fn(self._F, other._F), coordinate_map_key=self.coordinate_map_key,
                                coordinate_manager=self.coordinate_manager)
        out_key, maps = self.coordinate_manager.union_map([self.coordinate_map_key, other.coordinate_map_key])
        n_out = self.coordinate_manager.size(out_key)
        dev = self._F.device
        out = torch.zeros(n_out, self._F.shape[1], dtype=self._F.dtype, device=dev)
        out = out.index_add(0, maps[0][1].to(dev), self._F)
        contrib = fn(torch.zeros_like(other._F), other._F)   # 0 (op) b
        out = out.index_add(0, maps[1][1].to(dev), contrib)
        return SparseTensor(out, coordinate_map_key=out_key, coordinate_manager=self.coordinate_manager)

    def __add__(self, other):
        return self._binary(other, lambda a, b: a + b)

    def __sub__(self, other):
        return self._binary(other, lambda a, b: a - b)

    def __mul__(self, other):
        assert not isinstance(other, SparseTensor) or other.coordinate_map_key == self.coordinate_map_key
        return self._binary(other, lambda a, b: a * b)

    def __iadd__(self, other):
        return self.__add__(other)

    def detach(self):
        return SparseTensor(self._F.detach(), coordinate_map_key=self.coordinate_map_key,
                            coordinate_manager=self.coordinate_manager)

    def to(self, *a, **k):
        return SparseTensor(self._F.to(*a, **k), coordinate_map_key=self.coordinate_map_key,
                            coordinate_manager=self.coordinate_manager)


def _like(x: SparseTensor, feats: torch.Tensor) -> SparseTensor:
    return SparseTensor(feats, coordinate_map_key=x.coordinate_map_key, coordinate_manager=x.coordinate_manager)


# ----------------------------------------------------------------------------
# functional ops (the algorithm restated; autograd by composition)
# ----------------------------------------------------------------------------
def conv_apply(feats: torch.Tensor, nbr: torch.Tensor, weight: torch.Tensor) -> torch.Tensor:
    """out[o] = Σ_k feats[nbr[k,o]] @ weight[k]   — ME's per-offset gather → GEMM → scatter-add."""
    K, n_out = nbr.shape
    out = torch.zeros(n_out, weight.shape[-1], dtype=feats.dtype, device=feats.device)
    nbr = nbr.to(feats.device)
    for k in range(K):
        o = torch.nonzero(nbr[k] >= 0, as_tuple=True)[0]
        if o.numel() == 0:
            continue
        i = nbr[k][o]
        out = out.index_add(0, o, feats.index_select(0, i) @ weight[k])
    return out


def maxpool_apply(feats: torch.Tensor, nbr: torch.Tensor) -> torch.Tensor:
    K, n_out = nbr.shape
    nbr = nbr.to(feats.device)
    neg = torch.full((1, feats.shape[1]), float("-inf"), dtype=feats.dtype, device=feats.device)
    padded = torch.cat([feats, neg], 0)
    idx = torch.where(nbr >= 0, nbr, torch.full_like(nbr, feats.shape[0]))
    return padded[idx].max(0)[0]            # [K, n_out, C] → [n_out, C]


def scatter_max(src: torch.Tensor, index: torch.Tensor, dim: int = 0, out=None, dim_size=None):
    """torch_scatter.scatter_max(src[P,C], index[P], dim=0) → (out[N,C], argmax[N,C]);
    empty segments give 0 / P  (unet3d_sparse_v2.py:79)."""
    assert dim == 0 and src.ndim == 2
    n = int(index.max()) + 1 if dim_size is None else dim_size
    idx = index.view(-1, 1).expand_as(src)
    res = torch.full((n, src.shape[1]), float("-inf"), dtype=src.dtype, device=src.device)
    res = res.scatter_reduce(0, idx, src, reduce="amax", include_self=True)
    # argmax: first row attaining the max
    hit = src == res.gather(0, idx)
    rows = torch.arange(src.shape[0], device=src.device).view(-1, 1).expand_as(src)
    cand = torch.where(hit, rows, torch.full_like(rows, src.shape[0]))
    arg = torch.full((n, src.shape[1]), src.shape[0], dtype=torch.int64, device=src.device)
    arg = arg.scatter_reduce(0, idx, cand, reduce="amin", include_self=True)
    res = torch.where(torch.isinf(res) & (res < 0), torch.zeros_like(res), res)
    return res, arg


# ----------------------------------------------------------------------------
# modules (names / parameter names follow ME so state_dicts line up)
# ----------------------------------------------------------------------------
class MinkowskiModuleBase(nn.Module):
    pass


class _ConvBase(MinkowskiModuleBase):
    def __init__(self, in_channels, out_channels, kernel_size=-1, stride=1, dilation=1, bias=False,
                 kernel_generator=None, is_transpose=False, expand_coordinates=False,
                 convolution_mode=None, dimension=3):
        super().__init__()
        assert dimension == 3
        self.in_channels, self.out_channels = in_channels, out_channels
        self.kernel_size, self.stride, self.dilation = int(kernel_size), int(stride), int(dilation)
        self.is_transpose, self.expand_coordinates = is_transpose, expand_coordinates
        self.kernel_volume = self.kernel_size ** 3
        self.use_mm = self.kernel_volume == 1 and self.stride == 1
        shape = (in_channels, out_channels) if self.use_mm else (self.kernel_volume, in_channels, out_channels)
        self.kernel = nn.Parameter(torch.empty(*shape))
        self.bias = nn.Parameter(torch.empty(1, out_channels)) if bias else None
        self.reset_parameters()

    def reset_parameters(self):
        with torch.no_grad():
            n = (self.out_channels if self.is_transpose else self.in_channels) * self.kernel_volume
            stdv = 1.0 / np.sqrt(n)
            self.kernel.uniform_(-stdv, stdv)
            if self.bias is not None:
                self.bias.uniform_(-stdv, stdv)

    def forward(self, x: SparseTensor) -> SparseTensor:
        assert isinstance(x, SparseTensor)
        assert x.F.shape[1] == self.in_channels, \
            f"Channel size mismatch {x.F.shape[1]} != {self.in_channels}"
        cm, in_key = x.coordinate_manager, x.coordinate_map_key
        if self.use_mm:
            out_f, out_key = x.F @ self.kernel, in_key
        else:
            if self.is_transpose:
                assert self.expand_coordinates, "oracle: only generative transposed conv is on the path"
                out_key = cm.generate(in_key, self.kernel_size, self.stride)
            elif self.stride > 1:
                out_key = cm.stride(in_key, self.stride)
            else:
                out_key = in_key
            nbr = cm.kernel_map(in_key, out_key, self.kernel_size, self.dilation, self.is_transpose)
            out_f = conv_apply(x.F, nbr, self.kernel)
        if self.bias is not None:
            out_f = out_f + self.bias
        return SparseTensor(out_f, coordinate_map_key=out_key, coordinate_manager=cm)


class MinkowskiConvolution(_ConvBase):
    def __init__(self, in_channels, out_channels, kernel_size=-1, stride=1, dilation=1, bias=False,
                 kernel_generator=None, expand_coordinates=False, convolution_mode=None, dimension=None):
        super().__init__(in_channels, out_channels, kernel_size, stride, dilation, bias,
                         is_transpose=False, expand_coordinates=expand_coordinates, dimension=dimension)


class MinkowskiConvolutionTranspose(_ConvBase):
    def __init__(self, in_channels, out_channels, kernel_size=-1, stride=1, dilation=1, bias=False,
                 kernel_generator=None, expand_coordinates=False, convolution_mode=None, dimension=None):
        super().__init__(in_channels, out_channels, kernel_size, stride, dilation, bias,
                         is_transpose=True, expand_coordinates=expand_coordinates, dimension=dimension)


class MinkowskiBatchNorm(MinkowskiModuleBase):
    def __init__(self, num_features, eps=1e-5, momentum=0.1, affine=True, track_running_stats=True):
        super().__init__()
        self.bn = nn.BatchNorm1d(num_features, eps=eps, momentum=momentum, affine=affine,
                                 track_running_stats=track_running_stats)

    def forward(self, x):
        return _like(x, self.bn(x.F))


class MinkowskiSyncBatchNorm(MinkowskiBatchNorm):
    def __init__(self, num_features, eps=1e-5, momentum=0.1, affine=True, track_running_stats=True,
                 process_group=None):
        MinkowskiModuleBase.__init__(self)
        self.bn = nn.SyncBatchNorm(num_features, eps=eps, momentum=momentum, affine=affine,
                                   track_running_stats=track_running_stats, process_group=process_group)

    @classmethod
    def convert_sync_batchnorm(cls, module, process_group=None):
        out = module
        if isinstance(module, MinkowskiBatchNorm) and not isinstance(module, MinkowskiSyncBatchNorm):
            out = MinkowskiSyncBatchNorm(module.bn.num_features, module.bn.eps, module.bn.momentum,
                                         module.bn.affine, module.bn.track_running_stats, process_group)
            if module.bn.affine:
                with torch.no_grad():
                    out.bn.weight = module.bn.weight
                    out.bn.bias = module.bn.bias
            out.bn.running_mean = module.bn.running_mean
            out.bn.running_var = module.bn.running_var
            out.bn.num_batches_tracked = module.bn.num_batches_tracked
            return out
        for name, child in module.named_children():
            out.add_module(name, cls.convert_sync_batchnorm(child, process_group))
        return out


def _elementwise(name, torch_cls):
    class _M(MinkowskiModuleBase):
        def __init__(self, *a, **k):
            super().__init__()
            self.module = torch_cls(*a, **k)

        def forward(self, x):
            return _like(x, self.module(x.F))
    _M.__name__ = _M.__qualname__ = name
    return _M


MinkowskiReLU = _elementwise("MinkowskiReLU", nn.ReLU)
MinkowskiLeakyReLU = _elementwise("MinkowskiLeakyReLU", nn.LeakyReLU)
MinkowskiSigmoid = _elementwise("MinkowskiSigmoid", nn.Sigmoid)
MinkowskiSoftmax = _elementwise("MinkowskiSoftmax", nn.Softmax)
MinkowskiDropout = _elementwise("MinkowskiDropout", nn.Dropout)
MinkowskiGELU = _elementwise("MinkowskiGELU", nn.GELU)
MinkowskiTanh = _elementwise("MinkowskiTanh", nn.Tanh)


class MinkowskiLinear(MinkowskiModuleBase):
    def __init__(self, in_features, out_features, bias=True):
        super().__init__()
        self.linear = nn.Linear(in_features, out_features, bias=bias)

    def forward(self, x):
        return _like(x, self.linear(x.F))


class MinkowskiPruning(MinkowskiModuleBase):
    def forward(self, x: SparseTensor, mask: torch.Tensor) -> SparseTensor:
        assert isinstance(mask, torch.Tensor) and mask.dtype == torch.bool
        if mask.ndim != 1 or mask.shape[0] != x.F.shape[0]:
            raise RuntimeError(f"MinkowskiPruning: mask length {tuple(mask.shape)} != rows {x.F.shape[0]}")
        rows = torch.nonzero(mask, as_tuple=True)[0]
        cm = x.coordinate_manager
        coords = cm.get_coordinates(x.coordinate_map_key)[rows.cpu()]
        key = cm._register(coords, x.coordinate_map_key.tensor_stride)
        return SparseTensor(x.F[rows], coordinate_map_key=key, coordinate_manager=cm)


class MinkowskiMaxPooling(MinkowskiModuleBase):
    def __init__(self, kernel_size, stride=1, dilation=1, kernel_generator=None, dimension=None):
        super().__init__()
        self.kernel_size, self.stride, self.dilation = int(kernel_size), int(stride), int(dilation)

    def forward(self, x: SparseTensor) -> SparseTensor:
        cm, in_key = x.coordinate_manager, x.coordinate_map_key
        out_key = cm.stride(in_key, self.stride) if self.stride > 1 else in_key
        nbr = cm.kernel_map(in_key, out_key, self.kernel_size, self.dilation, False)
        return SparseTensor(maxpool_apply(x.F, nbr), coordinate_map_key=out_key, coordinate_manager=cm)


class _NotOnPath(MinkowskiModuleBase):
    """Symbols that only appear in never-instantiated reference classes."""

    def __init__(self, *a, **k):
        raise NotImplementedError(f"{type(self).__name__} is not on PaSCo's live path")


class MinkowskiGlobalPooling(_NotOnPath):
    pass


class MinkowskiBroadcastMultiplication(_NotOnPath):
    pass


class MinkowskiChannelwiseConvolution(_NotOnPath):
    pass


class MinkowskiPoolingTranspose(_NotOnPath):
    pass


class MinkowskiGlobalMaxPooling(_NotOnPath):
    pass


def cat(*tensors):
    if len(tensors) == 1 and isinstance(tensors[0], (list, tuple)):
        tensors = tensors[0]
    k = tensors[0].coordinate_map_key
    assert all(t.coordinate_map_key == k for t in tensors)
    return _like(tensors[0], torch.cat([t.F for t in tensors], 1))


def to_sparse(x: torch.Tensor, format=None, coordinates=None, device=None) -> SparseTensor:
    """dense [B,C,X,Y,Z] → rows where |x|.sum(C) != 0, torch.where order (b,x,y,z)."""
    assert x.ndim == 5
    if coordinates is None:
        nz = torch.where(x.abs().sum(1) != 0)
        coordinates = torch.stack(nz, 1).int()
    c = coordinates.long()
    feats = x[c[:, 0], :, c[:, 1], c[:, 2], c[:, 3]]
    return SparseTensor(feats, coordinates=coordinates.int())


class _Utils:
    @staticmethod
    def batched_coordinates(coords: List[torch.Tensor], dtype=torch.int32, device=None):
        out = []
        for b, c in enumerate(coords):
            c = torch.as_tensor(c)
            bc = torch.full((c.shape[0], 1), b, dtype=c.dtype, device=c.device)
            out.append(torch.cat([bc, c], 1))
        res = torch.cat(out, 0).to(dtype) if out else torch.zeros(0, 4, dtype=dtype)
        return res if device is None else res.to(device)

    @staticmethod
    def sparse_collate(coords, feats, labels=None, dtype=torch.int32, device=None):
        bc = _Utils.batched_coordinates(coords, dtype, device)
        f = torch.cat([torch.as_tensor(x) for x in feats], 0)
        if labels is None:
            return bc, f
        return bc, f, torch.cat([torch.as_tensor(x) for x in labels], 0)


utils = _Utils()
