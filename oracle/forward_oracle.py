"""ORACLE (test infrastructure, not product code): CPU fp32 restatement of DUSt3R's pairwise forward.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this file.  The product path (dust3r_b200/) never does.

Parity status: pinned.  tests/test_oracle_vs_reference.py runs this restatement against the unmodified
reference (imported from /root/reference when it is mounted) on seeded weights/inputs, and
tests/golden/*.npz hold outputs of the reference itself (made by tests/golden/make_golden.py).

The function takes a plain state dict (reference key names) and returns every stage the parity tests
look at.  It is written functionally (no nn.Module tree) and follows:

  patch embed + positions ....... dust3r/patch_embed.py:19-29, croco/models/blocks.py:195-207
  encoder block .................. croco/models/blocks.py:81-130
  2D RoPE ........................ croco/models/pos_embed.py:113-157 (== curope/kernels.cu:17-82 math)
  symmetrised encode ............. dust3r/model.py:142-170
  decoder (twin, cross-attn) ..... dust3r/model.py:172-191, croco/models/blocks.py:132-191
  DPT head ....................... dust3r/heads/dpt_head.py:34-65, croco/models/dpt_block.py:82-262
  linear head .................... dust3r/heads/linear_head.py:30-41
  postprocess .................... dust3r/heads/postprocess.py:10-58
"""
from __future__ import annotations

import math
import torch
import torch.nn.functional as F

LN_EPS = 1e-6  # croco/models/croco.py:33 (partial(nn.LayerNorm, eps=1e-6))

# Calibration mode (tests only): with OPERAND_DTYPE = torch.bfloat16 every contraction (linear, convolution, Q K^T, P V) sees
# operands rounded to bf16 and accumulates in fp32 -- what ANY bf16-operand / fp32-accumulate implementation computes up to
# summation order.  The distance between this and the plain fp32 oracle is the error the product is entitled to; the
# distance between the product and this mode is what is left for implementation error.  Default None = the reference's fp32.
OPERAND_DTYPE = None


class operand_rounding:
    def __init__(self, dtype):
        self.dtype = dtype

    def __enter__(self):
        global OPERAND_DTYPE
        self.prev, OPERAND_DTYPE = OPERAND_DTYPE, self.dtype

    def __exit__(self, *a):
        global OPERAND_DTYPE
        OPERAND_DTYPE = self.prev


def _r(t):
    return t if OPERAND_DTYPE is None or t is None else t.to(OPERAND_DTYPE).to(torch.float32)


def _ln(x, sd, prefix):
    return F.layer_norm(x, (x.shape[-1],), sd[prefix + '.weight'], sd[prefix + '.bias'], LN_EPS)


def _lin(x, sd, prefix):
    return F.linear(_r(x), _r(sd[prefix + '.weight']), sd[prefix + '.bias'])


def positions(B, h, w):
    """(B, h*w, 2) int64, [:, :, 0] = row (y), [:, :, 1] = column (x).  blocks.py:201-207."""
    ys, xs = torch.meshgrid(torch.arange(h), torch.arange(w), indexing='ij')
    pos = torch.stack((ys.reshape(-1), xs.reshape(-1)), dim=-1)
    return pos[None].expand(B, -1, -1).clone()


def rope_tables(hd, npos, base):
    """cos/sin (npos, hd//4): angle[p, k] = p * base**(-k/(hd//4)).  pos_embed.py:121-131 with D=hd//2."""
    q = hd // 4
    inv_freq = 1.0 / (base ** (torch.arange(0, q).float() / q))
    ang = torch.arange(npos).float()[:, None] * inv_freq[None, :]
    return ang.cos(), ang.sin()


def rope2d(t, pos, base):
    """t: (B, H, N, hd) ; pos: (B, N, 2).  First half of hd rotated by y, second half by x; inside
    each half, element k pairs with k + hd/4 (rotate_half).  pos_embed.py:133-157."""
    B, H, N, hd = t.shape
    q = hd // 4
    cos, sin = rope_tables(hd, int(pos.max()) + 1, base)
    out = torch.empty_like(t)
    for half, col in ((0, 0), (1, 1)):
        c = cos[pos[:, :, col]][:, None]      # (B,1,N,q)
        s = sin[pos[:, :, col]][:, None]
        u = t[..., half * 2 * q: half * 2 * q + q]
        v = t[..., half * 2 * q + q: half * 2 * q + 2 * q]
        out[..., half * 2 * q: half * 2 * q + q] = u * c - v * s
        out[..., half * 2 * q + q: half * 2 * q + 2 * q] = v * c + u * s
    return out


def _attention(q, k, v, scale):
    attn = (_r(q) @ _r(k).transpose(-2, -1)) * scale
    attn = attn.softmax(dim=-1)
    return _r(attn) @ _r(v)


def self_attention(x, pos, sd, prefix, nh, base):
    B, N, C = x.shape
    hd = C // nh
    qkv = _lin(x, sd, prefix + '.qkv').reshape(B, N, 3, nh, hd).permute(2, 0, 3, 1, 4)
    q, k, v = qkv[0], qkv[1], qkv[2]
    q = rope2d(q, pos, base)
    k = rope2d(k, pos, base)
    o = _attention(q, k, v, hd ** -0.5).transpose(1, 2).reshape(B, N, C)
    return _lin(o, sd, prefix + '.proj')


def cross_attention(x, y, xpos, ypos, sd, prefix, nh, base):
    B, Nq, C = x.shape
    Nk = y.shape[1]
    hd = C // nh
    q = _lin(x, sd, prefix + '.projq').reshape(B, Nq, nh, hd).permute(0, 2, 1, 3)
    k = _lin(y, sd, prefix + '.projk').reshape(B, Nk, nh, hd).permute(0, 2, 1, 3)
    v = _lin(y, sd, prefix + '.projv').reshape(B, Nk, nh, hd).permute(0, 2, 1, 3)
    q = rope2d(q, xpos, base)
    k = rope2d(k, ypos, base)
    o = _attention(q, k, v, hd ** -0.5).transpose(1, 2).reshape(B, Nq, C)
    return _lin(o, sd, prefix + '.proj')


def mlp(x, sd, prefix):
    return _lin(F.gelu(_lin(x, sd, prefix + '.fc1')), sd, prefix + '.fc2')


def enc_block(x, pos, sd, prefix, nh, base):
    x = x + self_attention(_ln(x, sd, prefix + '.norm1'), pos, sd, prefix + '.attn', nh, base)
    x = x + mlp(_ln(x, sd, prefix + '.norm2'), sd, prefix + '.mlp')
    return x


def dec_block(x, y, xpos, ypos, sd, prefix, nh, base):
    x = x + self_attention(_ln(x, sd, prefix + '.norm1'), xpos, sd, prefix + '.attn', nh, base)
    y_ = _ln(y, sd, prefix + '.norm_y')
    x = x + cross_attention(_ln(x, sd, prefix + '.norm2'), y_, xpos, ypos, sd, prefix + '.cross_attn', nh, base)
    x = x + mlp(_ln(x, sd, prefix + '.norm3'), sd, prefix + '.mlp')
    return x


def encode(img, sd, cfg, stages=None):
    """img (B,3,H,W) -> tokens (B,N,E), pos (B,N,2)."""
    B, _, H, W = img.shape
    p = cfg.patch_size
    assert H % p == 0 and W % p == 0
    x = F.conv2d(_r(img), _r(sd['patch_embed.proj.weight']), sd['patch_embed.proj.bias'], stride=p)
    x = x.flatten(2).transpose(1, 2)
    pos = positions(B, H // p, W // p)
    if stages is not None:
        stages['patch_embed'] = x
    for i in range(cfg.enc_depth):
        x = enc_block(x, pos, sd, f'enc_blocks.{i}', cfg.enc_num_heads, cfg.rope_freq)
        if stages is not None:
            stages[f'enc_block{i}'] = x
    x = _ln(x, sd, 'enc_norm')
    if stages is not None:
        stages['enc_norm'] = x
    return x, pos


def decode(f1, pos1, f2, pos2, sd, cfg, stages=None):
    """Returns the two 13-element (dec_depth+1) token lists the heads receive (model.py:172-191)."""
    outs1, outs2 = [f1], [f2]
    a = _lin(f1, sd, 'decoder_embed')
    b = _lin(f2, sd, 'decoder_embed')
    if stages is not None:
        stages['decoder_embed1'], stages['decoder_embed2'] = a, b
    for i in range(cfg.dec_depth):
        na = dec_block(a, b, pos1, pos2, sd, f'dec_blocks.{i}', cfg.dec_num_heads, cfg.rope_freq)
        nb = dec_block(b, a, pos2, pos1, sd, f'dec_blocks2.{i}', cfg.dec_num_heads, cfg.rope_freq)
        a, b = na, nb
        outs1.append(a)
        outs2.append(b)
        if stages is not None:
            stages[f'dec_block{i}_1'], stages[f'dec_block{i}_2'] = a, b
    outs1[-1] = _ln(outs1[-1], sd, 'dec_norm')
    outs2[-1] = _ln(outs2[-1], sd, 'dec_norm')
    return outs1, outs2


def _conv(x, sd, prefix, **kw):
    return F.conv2d(_r(x), _r(sd[prefix + '.weight']), sd.get(prefix + '.bias'), **kw)


def _rcu(x, sd, prefix):
    out = _conv(F.relu(x), sd, prefix + '.conv1', padding=1)
    out = _conv(F.relu(out), sd, prefix + '.conv2', padding=1)
    return out + x


def _fusion(sd, prefix, x, skip=None):
    if skip is not None:
        x = x + _rcu(skip, sd, prefix + '.resConfUnit1')
    x = _rcu(x, sd, prefix + '.resConfUnit2')
    x = F.interpolate(x, scale_factor=2, mode='bilinear', align_corners=True)
    return _conv(x, sd, prefix + '.out_conv')


def dpt_head(tokens, H, W, sd, prefix, cfg, stages=None, tag=''):
    """tokens: list of dec_depth+1 tensors (B,N,C).  Returns (B,4,H,W) before postprocess."""
    p = cfg.patch_size
    nh, nw = H // p, W // p
    layers = []
    for hook in cfg.dpt_hooks:
        t = tokens[hook]
        B = t.shape[0]
        layers.append(t.transpose(1, 2).reshape(B, t.shape[2], nh, nw))
    ap = prefix + '.act_postprocess'
    l0 = _conv(layers[0], sd, ap + '.0.0')
    l0 = F.conv_transpose2d(_r(l0), _r(sd[ap + '.0.1.weight']), sd[ap + '.0.1.bias'], stride=4)
    l1 = _conv(layers[1], sd, ap + '.1.0')
    l1 = F.conv_transpose2d(_r(l1), _r(sd[ap + '.1.1.weight']), sd[ap + '.1.1.bias'], stride=2)
    l2 = _conv(layers[2], sd, ap + '.2.0')
    l3 = _conv(layers[3], sd, ap + '.3.0')
    l3 = _conv(l3, sd, ap + '.3.1', stride=2, padding=1)
    ls = [l0, l1, l2, l3]
    ls = [_conv(l, sd, f'{prefix}.scratch.layer_rn.{k}', padding=1) for k, l in enumerate(ls)]
    sc = prefix + '.scratch'
    path4 = _fusion(sd, sc + '.refinenet4', ls[3])[:, :, :ls[2].shape[2], :ls[2].shape[3]]
    path3 = _fusion(sd, sc + '.refinenet3', path4, ls[2])
    path2 = _fusion(sd, sc + '.refinenet2', path3, ls[1])
    path1 = _fusion(sd, sc + '.refinenet1', path2, ls[0])
    if stages is not None:
        for k, l in enumerate(ls):
            stages[f'dpt{tag}_layer{k}'] = l
        stages[f'dpt{tag}_path4'], stages[f'dpt{tag}_path3'] = path4, path3
        stages[f'dpt{tag}_path2'], stages[f'dpt{tag}_path1'] = path2, path1
    out = _conv(path1, sd, prefix + '.head.0', padding=1)
    out = F.interpolate(out, scale_factor=2, mode='bilinear', align_corners=True)
    out = F.relu(_conv(out, sd, prefix + '.head.2', padding=1))
    out = _conv(out, sd, prefix + '.head.4')
    return out


def linear_head(tokens, H, W, sd, prefix, cfg):
    p = cfg.patch_size
    t = tokens[-1]
    B = t.shape[0]
    feat = _lin(t, sd, prefix + '.proj')
    feat = feat.transpose(-1, -2).reshape(B, -1, H // p, W // p)
    return F.pixel_shuffle(feat, p)


def postprocess(out, cfg):
    """(B,4,H,W) -> pts3d (B,H,W,3), conf (B,H,W).  postprocess.py:10-58 for the published modes."""
    fmap = out.permute(0, 2, 3, 1)
    mode, vmin, vmax = cfg.depth_mode
    assert vmin == -float('inf') and vmax == float('inf')
    xyz = fmap[..., 0:3]
    if mode == 'linear':
        pts = xyz
    else:
        d = xyz.norm(dim=-1, keepdim=True)
        xyz = xyz / d.clip(min=1e-8)
        if mode == 'square':
            pts = xyz * d.square()
        elif mode == 'exp':
            pts = xyz * torch.expm1(d)
        else:
            raise ValueError(mode)
    res = {'pts3d': pts}
    if cfg.conf_mode is not None:
        cmode, cmin, cmax = cfg.conf_mode
        x = fmap[..., 3]
        if cmode == 'exp':
            res['conf'] = cmin + x.exp().clip(max=cmax - cmin)
        elif cmode == 'sigmoid':
            res['conf'] = (cmax - cmin) * torch.sigmoid(x) + cmin
        else:
            raise ValueError(cmode)
    return res


def is_symmetrized(inst1, inst2):
    """dust3r/utils/misc.py:32-40."""
    if len(inst1) == len(inst2) == 1:
        return False
    ok = True
    for i in range(0, len(inst1), 2):
        ok = ok and (inst1[i] == inst2[i + 1]) and (inst1[i + 1] == inst2[i])
    return ok


@torch.no_grad()
def forward_oracle(sd, cfg, img1, img2, instance1=None, instance2=None, stages=None):
    """Full pair forward.  Returns (res1, res2) like AsymmetricCroCo3DStereo.forward (model.py:199-211).
    Pairs whose two images differ in size are encoded separately (model.py:147-151)."""
    sd = {k: v.float() for k, v in sd.items()}
    B = img1.shape[0]
    sym = instance1 is not None and is_symmetrized(instance1, instance2)
    if sym:
        e, pos = encode(torch.cat((img1[::2], img2[::2])), sd, cfg, stages)
        a, b = e.chunk(2)
        pa, pb = pos.chunk(2)
        f1 = torch.stack((a, b), 1).flatten(0, 1)
        f2 = torch.stack((b, a), 1).flatten(0, 1)
        pos1 = torch.stack((pa, pb), 1).flatten(0, 1)
        pos2 = torch.stack((pb, pa), 1).flatten(0, 1)
    elif img1.shape[-2:] != img2.shape[-2:]:
        # model.py:147-151 (_encode_image_pairs): images of different sizes are encoded separately
        f1, pos1 = encode(img1, sd, cfg, stages)
        f2, pos2 = encode(img2, sd, cfg, None)
    else:
        e, pos = encode(torch.cat((img1, img2)), sd, cfg, stages)
        f1, f2 = e.chunk(2)
        pos1, pos2 = pos.chunk(2)
    d1, d2 = decode(f1, pos1, f2, pos2, sd, cfg, stages)
    H1, W1 = img1.shape[-2:]
    H2, W2 = img2.shape[-2:]
    if cfg.head_type == 'dpt':
        o1 = dpt_head(d1, H1, W1, sd, 'downstream_head1.dpt', cfg, stages, '1')
        o2 = dpt_head(d2, H2, W2, sd, 'downstream_head2.dpt', cfg, stages, '2')
    else:
        o1 = linear_head(d1, H1, W1, sd, 'downstream_head1', cfg)
        o2 = linear_head(d2, H2, W2, sd, 'downstream_head2', cfg)
    if stages is not None:
        stages['head1_raw'], stages['head2_raw'] = o1, o2
    r1 = postprocess(o1, cfg)
    r2 = postprocess(o2, cfg)
    r2['pts3d_in_other_view'] = r2.pop('pts3d')
    return r1, r2
