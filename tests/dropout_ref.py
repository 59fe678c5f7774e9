"""Host replica of gslora-b200's counter-based dropout masks (csrc/gsl_common.cuh drop_hash / gsl_engine.cu site_seed), so the
oracle can be run with exactly the masks the engine draws (torch's Philox stream cannot be matched bit-for-bit)."""
import torch

M32 = 0xFFFFFFFF


def drop_hash(pair: torch.Tensor, seed: int) -> torch.Tensor:
    h = ((pair * 0x9E3779B1) & M32) ^ seed
    h = h ^ (h >> 16); h = (h * 0x85EBCA6B) & M32
    h = h ^ (h >> 13); h = (h * 0xC2B2AE35) & M32
    return h ^ (h >> 16)


def drop_bits(pair: torch.Tensor, seed: int) -> torch.Tensor:
    """csrc/gsl_common.cuh drop_bits: the per-pair mask hash (two multiply / xor-shift rounds)."""
    h = ((pair ^ seed) * 0x9E3779B1) & M32
    h = h ^ (h >> 15)
    return (h * 0x85EBCA77) & M32


def site_seed(base: int, block: int, site: int) -> int:
    s = (base & M32) ^ ((base >> 32) & M32)
    return int(drop_hash(torch.tensor([block * 4 + site + 1], dtype=torch.int64), s).item())


def keep_mask(rows: int, cols: int, p: float, seed: int) -> torch.Tensor:
    """[rows, cols] float mask with values 0 or 1/(1-p); element index e = row * cols + col, pair e >> 1, 15 bits per element."""
    if p <= 0:
        return torch.ones(rows, cols)
    e = torch.arange(rows * cols, dtype=torch.int64)
    h = drop_bits(e >> 1, seed)
    bits = torch.where((e & 1) == 1, h >> 16, h) & 0x7FFF
    thresh = int(p * 32768.0 + 0.5)
    return ((bits >= thresh).float() / (1.0 - p)).view(rows, cols)


def engine_masks(cfg, B: int, base_seed: int, p: float, p_emb: float):
    """The four dropout sites of ViT_face for a batch of B images, keyed like oracle.vit_oracle.vit_embed(masks=...)."""
    N, D, H, L = cfg.tokens, cfg.dim, cfg.mlp_dim, cfg.depth
    m = {"emb": keep_mask(B * N, D, p_emb, site_seed(base_seed, L, 0)).view(B, N, D)}
    for i in range(L):
        if i < L - 1:
            m[("attn", i)] = keep_mask(B * N, D, p, site_seed(base_seed, i, 1)).view(B, N, D)
            m[("gelu", i)] = keep_mask(B * N, H, p, site_seed(base_seed, i, 2)).view(B, N, H)
            m[("ffn", i)] = keep_mask(B * N, D, p, site_seed(base_seed, i, 3)).view(B, N, D)
        else:
            # last block: the engine only computes the B cls rows (compact row index b); the other tokens of this block never
            # reach the loss, so their masks are irrelevant (ones)
            for name, site, width in (("attn", 1, D), ("gelu", 2, H), ("ffn", 3, D)):
                full = torch.ones(B, N, width)
                full[:, 0, :] = keep_mask(B, width, p, site_seed(base_seed, i, site))
                m[(name, i)] = full
    return m
