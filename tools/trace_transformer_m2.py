"""Diagnostic: MIMO (M=2, zero-padded) MaskTransformer on CUDA vs a CPU restatement validated against the reference."""
import os, sys, copy
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle")]
import torch
import torch.nn.functional as F
import me_oracle as OR
from pasco_b200 import build
build.build()
from pasco_b200 import me as ME, ops
from pasco_b200.net3d import MaskTransformer, sine_position_encoding

torch.manual_seed(0)
M = 2
cpu = MaskTransformer([256, 128, 64], 20, 384, 100, 8, 1024, 64, M)
gpu = copy.deepcopy(cpu).cuda()
g = torch.Generator().manual_seed(1)


def mk(scale, C, p):
    occ = torch.rand(32 // scale, 32 // scale, 8 // scale, generator=g) < p
    c = torch.nonzero(occ).int() * scale
    bc = OR.utils.batched_coordinates([c])
    return torch.randn(bc.shape[0], C, generator=g), bc


subs = [{4: mk(4, 256, 0.6), 2: mk(2, 128, 0.5), 1: mk(1, 64, 0.3)}, {4: mk(4, 256, 0.4), 2: mk(2, 128, 0.6), 1: mk(1, 64, 0.25)}]
maxlen = {s: max(sub[s][0].shape[0] for sub in subs) for s in (4, 2, 1)}


def my_forward(tf, feats, subnet):
    nq = tf.num_queries
    o = tf.query_feat.weight[subnet * nq:(subnet + 1) * nq].unsqueeze(0)
    qpos = tf.query_embed.weight[subnet * nq:(subnet + 1) * nq].unsqueeze(0)
    n_pos = tf.hidden_dim // 3
    F1, C1 = feats[1]
    uidx, _ = OR.unique_first(C1)
    C1u = C1[uidx]
    pos1 = sine_position_encoding(C1[:, 1:], n_pos)
    vf = F.linear(F1, tf.mask_feat_proj.weight, tf.mask_feat_proj.bias) + pos1

    def heads(o):
        d = tf.decoder_norm(o)
        return tf.class_embed(d), vf @ tf.mask_embed(d)[0].t()
    cls, msk = heads(o)
    classes, masks, ams = [cls], [msk], []
    for i, s in enumerate(tf.src_scales):
        Fs, Cs = feats[s]
        kv = F.linear(Fs, tf.input_projs[i].weight, tf.input_projs[i].bias) + (pos1 if s == 1 else sine_position_encoding(Cs[:, 1:], n_pos))
        keep = OR.SparseTensor((msk.detach() > 0).float()[uidx], C1u)
        if s != 1:
            keep = OR.MinkowskiMaxPooling(kernel_size=s, stride=s, dimension=3)(keep)
        rows = OR.lookup(keep.C, Cs)
        at = torch.where(rows[:, None] >= 0, keep.F[rows.clamp(min=0)], torch.zeros(1))
        m = (at == 0).t().contiguous()
        m[m.all(1)] = False
        ams.append(m)
        ca = tf.transformer_cross_attention_layers[i]
        mha = ca.multihead_attn
        d = 384
        qn = ca.norm(o)
        wq, wk, wv = mha.in_proj_weight.split(d, 0)
        bq, bk, bv = mha.in_proj_bias.split(d, 0)
        Q, K, V = F.linear(qn + qpos, wq, bq)[0], F.linear(kv, wk, bk), F.linear(kv, wv, bv)
        hv = lambda t: t.view(t.shape[0], 8, 48).transpose(0, 1)
        S = torch.bmm(hv(Q) * 48 ** -0.5, hv(K).transpose(1, 2)).masked_fill(m.unsqueeze(0), float("-inf"))
        oo = torch.bmm(torch.softmax(S, -1), hv(V)).transpose(0, 1).reshape(1, -1, d)
        o = qn + mha.out_proj(oo)
        o = tf.transformer_self_attention_layers[i](o, qpos)
        o = tf.transformer_ffn_layers[i](o)
        cls, msk = heads(o)
        classes.append(cls)
        masks.append(msk)
    return classes, masks, ams


gpu_masks = []
orig = MaskTransformer.attn_mask


def attn_mask_gpu(self, *a):
    m = orig(self, *a)
    gpu_masks.append(m.cpu())
    return m


MaskTransformer.attn_mask = attn_mask_gpu
for m, sub in enumerate(subs):
    feats = {}
    for s, (Ft, Ct) in sub.items():
        pad = maxlen[s] - Ft.shape[0]
        c = Ct.clone()
        c[:, 0] = 0
        feats[s] = (F.pad(Ft, (0, 0, 0, pad)), F.pad(c, (0, 0, 0, pad)))
    with torch.no_grad():
        c_cls, c_msk, c_am = my_forward(cpu, feats, m)
        gpu_masks.clear()
        g_cls, g_msk = gpu({s: (a.cuda(), b.cuda()) for s, (a, b) in feats.items()}, m)
    for i in range(4):
        print("subnet", m, "level", i, "mask-logit err", float((g_msk[i].cpu() - c_msk[i]).abs().max()), "class err",
              float((g_cls[i].cpu() - c_cls[i]).abs().max()))
    for i, (a, b) in enumerate(zip(gpu_masks, c_am)):
        print("   attn mask", i, tuple(a.shape), "mismatches", int((a != b).sum()))
