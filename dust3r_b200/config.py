"""Model configuration and state-dict layout of the pairwise-forward hot path.

The key names / shapes reproduce the reference checkpoint layout so real DUSt3R
checkpoints load unchanged (reference: dust3r/model.py:58-126, croco/models/croco.py:22-109,
croco/models/blocks.py:81-191, croco/models/dpt_block.py:270-410, dust3r/heads/linear_head.py:12-41).
tests/test_state_dict_layout.py pins this against key lists dumped from the reference.
"""
from __future__ import annotations

from collections import OrderedDict
from dataclasses import dataclass, field, asdict
from typing import Tuple

inf = float('inf')


@dataclass
class ModelConfig:
    img_size: Tuple[int, int] = (224, 224)
    patch_size: int = 16
    enc_embed_dim: int = 1024
    enc_depth: int = 24
    enc_num_heads: int = 16
    dec_embed_dim: int = 768
    dec_depth: int = 12
    dec_num_heads: int = 12
    mlp_ratio: int = 4
    pos_embed: str = 'RoPE100'
    head_type: str = 'linear'          # 'linear' | 'dpt'
    output_mode: str = 'pts3d'
    depth_mode: tuple = ('exp', -inf, inf)
    conf_mode: tuple = ('exp', 1, inf)
    landscape_only: bool = True
    norm_im2_in_dec: bool = True
    # DPT head constants (dust3r/heads/dpt_head.py:95-115)
    dpt_feature_dim: int = 256
    dpt_layer_dims: Tuple[int, int, int, int] = (96, 192, 384, 768)

    @property
    def rope_freq(self) -> float:
        assert self.pos_embed.startswith('RoPE'), 'only RoPE positional embedding is on the DUSt3R path'
        return float(self.pos_embed[len('RoPE'):])

    @property
    def has_conf(self) -> bool:
        return bool(self.conf_mode)

    @property
    def dpt_hooks(self):
        l2 = self.dec_depth
        return [0, l2 * 2 // 4, l2 * 3 // 4, l2]

    @property
    def dpt_dim_tokens(self):
        return [self.enc_embed_dim, self.dec_embed_dim, self.dec_embed_dim, self.dec_embed_dim]

    def to_dict(self):
        return asdict(self)


# The two published architectures (README.md:99-103 of the reference; ctor strings README.md:364,388
# after load_model's rewriting, dust3r/model.py:27-43).
def vitl_512_dpt(**kw) -> ModelConfig:
    return ModelConfig(img_size=(512, 512), head_type='dpt', landscape_only=False, **kw)


def vitl_224_linear(**kw) -> ModelConfig:
    return ModelConfig(img_size=(224, 224), head_type='linear', landscape_only=False, **kw)


def _block_keys(prefix, dim, hidden, out):
    out[f'{prefix}.norm1.weight'] = (dim,)
    out[f'{prefix}.norm1.bias'] = (dim,)
    out[f'{prefix}.attn.qkv.weight'] = (3 * dim, dim)
    out[f'{prefix}.attn.qkv.bias'] = (3 * dim,)
    out[f'{prefix}.attn.proj.weight'] = (dim, dim)
    out[f'{prefix}.attn.proj.bias'] = (dim,)
    out[f'{prefix}.norm2.weight'] = (dim,)
    out[f'{prefix}.norm2.bias'] = (dim,)
    out[f'{prefix}.mlp.fc1.weight'] = (hidden, dim)
    out[f'{prefix}.mlp.fc1.bias'] = (hidden,)
    out[f'{prefix}.mlp.fc2.weight'] = (dim, hidden)
    out[f'{prefix}.mlp.fc2.bias'] = (dim,)


def _dec_block_keys(prefix, dim, hidden, out, norm_mem=True):
    out[f'{prefix}.norm1.weight'] = (dim,)
    out[f'{prefix}.norm1.bias'] = (dim,)
    out[f'{prefix}.attn.qkv.weight'] = (3 * dim, dim)
    out[f'{prefix}.attn.qkv.bias'] = (3 * dim,)
    out[f'{prefix}.attn.proj.weight'] = (dim, dim)
    out[f'{prefix}.attn.proj.bias'] = (dim,)
    for p in ('projq', 'projk', 'projv', 'proj'):
        out[f'{prefix}.cross_attn.{p}.weight'] = (dim, dim)
        out[f'{prefix}.cross_attn.{p}.bias'] = (dim,)
    out[f'{prefix}.norm2.weight'] = (dim,)
    out[f'{prefix}.norm2.bias'] = (dim,)
    out[f'{prefix}.norm3.weight'] = (dim,)
    out[f'{prefix}.norm3.bias'] = (dim,)
    out[f'{prefix}.mlp.fc1.weight'] = (hidden, dim)
    out[f'{prefix}.mlp.fc1.bias'] = (hidden,)
    out[f'{prefix}.mlp.fc2.weight'] = (dim, hidden)
    out[f'{prefix}.mlp.fc2.bias'] = (dim,)
    if norm_mem:
        out[f'{prefix}.norm_y.weight'] = (dim,)
        out[f'{prefix}.norm_y.bias'] = (dim,)


def _dpt_keys(prefix, cfg: ModelConfig, out):
    fd = cfg.dpt_feature_dim
    ld = cfg.dpt_layer_dims
    nch = 3 + int(cfg.has_conf)
    for k in range(4):
        out[f'{prefix}.scratch.layer{k + 1}_rn.weight'] = (fd, ld[k], 3, 3)
    for k in range(4):  # same storage as layer{k+1}_rn, both names are serialised
        out[f'{prefix}.scratch.layer_rn.{k}.weight'] = (fd, ld[k], 3, 3)
    for r in (1, 2, 3, 4):
        out[f'{prefix}.scratch.refinenet{r}.out_conv.weight'] = (fd, fd, 1, 1)
        out[f'{prefix}.scratch.refinenet{r}.out_conv.bias'] = (fd,)
        for u in (1, 2):
            for c in (1, 2):
                out[f'{prefix}.scratch.refinenet{r}.resConfUnit{u}.conv{c}.weight'] = (fd, fd, 3, 3)
                out[f'{prefix}.scratch.refinenet{r}.resConfUnit{u}.conv{c}.bias'] = (fd,)
    out[f'{prefix}.head.0.weight'] = (fd // 2, fd, 3, 3)
    out[f'{prefix}.head.0.bias'] = (fd // 2,)
    out[f'{prefix}.head.2.weight'] = (fd // 2, fd // 2, 3, 3)
    out[f'{prefix}.head.2.bias'] = (fd // 2,)
    out[f'{prefix}.head.4.weight'] = (nch, fd // 2, 1, 1)
    out[f'{prefix}.head.4.bias'] = (nch,)
    dt = cfg.dpt_dim_tokens
    out[f'{prefix}.act_postprocess.0.0.weight'] = (ld[0], dt[0], 1, 1)
    out[f'{prefix}.act_postprocess.0.0.bias'] = (ld[0],)
    out[f'{prefix}.act_postprocess.0.1.weight'] = (ld[0], ld[0], 4, 4)
    out[f'{prefix}.act_postprocess.0.1.bias'] = (ld[0],)
    out[f'{prefix}.act_postprocess.1.0.weight'] = (ld[1], dt[1], 1, 1)
    out[f'{prefix}.act_postprocess.1.0.bias'] = (ld[1],)
    out[f'{prefix}.act_postprocess.1.1.weight'] = (ld[1], ld[1], 2, 2)
    out[f'{prefix}.act_postprocess.1.1.bias'] = (ld[1],)
    out[f'{prefix}.act_postprocess.2.0.weight'] = (ld[2], dt[2], 1, 1)
    out[f'{prefix}.act_postprocess.2.0.bias'] = (ld[2],)
    out[f'{prefix}.act_postprocess.3.0.weight'] = (ld[3], dt[3], 1, 1)
    out[f'{prefix}.act_postprocess.3.0.bias'] = (ld[3],)
    out[f'{prefix}.act_postprocess.3.1.weight'] = (ld[3], ld[3], 3, 3)
    out[f'{prefix}.act_postprocess.3.1.bias'] = (ld[3],)


def state_dict_spec(cfg: ModelConfig) -> 'OrderedDict[str, tuple]':
    """key -> shape, in the order the reference module tree serialises them."""
    out = OrderedDict()
    E, D, P = cfg.enc_embed_dim, cfg.dec_embed_dim, cfg.patch_size
    out['mask_token'] = (1, 1, D)
    out['patch_embed.proj.weight'] = (E, 3, P, P)
    out['patch_embed.proj.bias'] = (E,)
    for i in range(cfg.enc_depth):
        _block_keys(f'enc_blocks.{i}', E, int(E * cfg.mlp_ratio), out)
    out['enc_norm.weight'] = (E,)
    out['enc_norm.bias'] = (E,)
    out['decoder_embed.weight'] = (D, E)
    out['decoder_embed.bias'] = (D,)
    for i in range(cfg.dec_depth):
        _dec_block_keys(f'dec_blocks.{i}', D, int(D * cfg.mlp_ratio), out, cfg.norm_im2_in_dec)
    out['dec_norm.weight'] = (D,)
    out['dec_norm.bias'] = (D,)
    for i in range(cfg.dec_depth):
        _dec_block_keys(f'dec_blocks2.{i}', D, int(D * cfg.mlp_ratio), out, cfg.norm_im2_in_dec)
    for h in (1, 2):
        if cfg.head_type == 'dpt':
            _dpt_keys(f'downstream_head{h}.dpt', cfg, out)
        elif cfg.head_type == 'linear':
            nch = 3 + int(cfg.has_conf)
            out[f'downstream_head{h}.proj.weight'] = (nch * P * P, D)
            out[f'downstream_head{h}.proj.bias'] = (nch * P * P,)
        else:
            raise NotImplementedError(f'unexpected head_type={cfg.head_type!r}')
    return out
