"""One-process-per-GPU sharding of the denoise step over the NVLink/NVSwitch domain of one box.

The reference has no live multi-GPU path for one job (one Ray actor per device, whole
jobs only); this module introduces the partitioning the north star asks for, in the two places where the
path really shards:

* **CFG pair** -- the conditional and unconditional forwards of a step (engine/wan/shared/__init__.py:548-563)
  are independent given the same latents: ranks are laid out as ``cfg_size x sp_size`` (cfg-major), branch
  ``cfg_rank`` runs on its ``sp_size`` ranks, and the two [B,16,F,H,W] bf16 predictions are exchanged with ONE
  all-gather per step inside the pair {r, r + sp_size}.
* **token shards** -- inside a forward everything except the self-attention core is token-independent, so the
  token axis is split into ``sp_size`` contiguous shards; around the global self-attention the shards are
  re-partitioned from tokens to heads and back (Ulysses): all-to-all #1 sends each rank the q|k|v columns of
  its ``heads / sp_size`` heads for all tokens, all-to-all #2 returns the attention output to token shards.
  Results are identical to the unsharded forward (no windowing / approximation).
* **decoded frames** -- VAE spatial tiles are dealt round-robin over all ranks and reassembled with a single
  all-gather (``allgather_frames``).

Plumbing is ``torch.distributed`` (NCCL on GPUs, gloo in the CPU tests); all timing of multi-GPU runs is done
on the device by the caller.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Sequence, List, Optional, Tuple

import torch
import torch.distributed as dist


@dataclass
class ParallelContext:
    rank: int = 0
    world_size: int = 1
    cfg_size: int = 1
    sp_size: int = 1
    cfg_group: Optional[object] = None   # process group of my CFG pair
    sp_group: Optional[object] = None    # process group of my token-shard peers
    world_group: Optional[object] = None
    #: fuse the Ulysses exchange into the kernels over NVLink peer memory (PeerExchange) instead of NCCL all-to-all
    use_p2p: bool = False
    _p2p: Optional[dict] = None

    def peer_exchange(self, tokens_total: int, heads: int, head_dim: int, device) -> "PeerExchange":
        """Symmetric buffers for one (tokens, heads) shape; allocated and rendezvoused once (collective!)."""
        if self._p2p is None:
            self._p2p = {}
        key = (tokens_total, heads, head_dim)
        if key not in self._p2p:
            self._p2p[key] = PeerExchange(self, tokens_total, heads, head_dim, device)
        return self._p2p[key]

    def joint_peer_exchange(self, n_img_total: int, n_txt: int, heads: int, head_dim: int, img_first: bool,
                            device) -> "JointPeerExchange":
        """Symmetric buffers of the dual-stream (joint sequence) exchange for one shape; collective on first use."""
        if self._p2p is None:
            self._p2p = {}
        key = ("joint", n_img_total, n_txt, heads, head_dim, img_first)
        if key not in self._p2p:
            self._p2p[key] = JointPeerExchange(self, n_img_total, n_txt, heads, head_dim, img_first, device)
        return self._p2p[key]

    # ------------------------------------------------------------------------------------ construction
    @classmethod
    def single(cls) -> "ParallelContext":
        return cls()

    @classmethod
    def create(cls, use_cfg: bool = True, world_group=None, use_p2p: bool = False) -> "ParallelContext":
        """Build the cfg x sp layout for the initialised default process group.  Every rank must call this
        (``dist.new_group`` is collective)."""
        if not dist.is_available() or not dist.is_initialized():
            return cls.single()
        world, rank = dist.get_world_size(), dist.get_rank()
        cfg_size, sp_size = plan_layout(world, use_cfg)
        cfg_group = sp_group = None
        # sp groups: contiguous blocks of sp_size ranks; cfg groups: {r, r + sp_size}
        for c in range(cfg_size):
            ranks = list(range(c * sp_size, (c + 1) * sp_size))
            g = dist.new_group(ranks)
            if rank in ranks:
                sp_group = g
        for s in range(sp_size):
            ranks = [s + c * sp_size for c in range(cfg_size)]
            g = dist.new_group(ranks)
            if rank in ranks:
                cfg_group = g
        return cls(rank=rank, world_size=world, cfg_size=cfg_size, sp_size=sp_size, cfg_group=cfg_group,
                   sp_group=sp_group, world_group=world_group, use_p2p=use_p2p)

    @property
    def cfg_rank(self) -> int:
        return self.rank // self.sp_size

    @property
    def sp_rank(self) -> int:
        return self.rank % self.sp_size

    # ------------------------------------------------------------------------------------ CFG exchange
    def exchange_cfg(self, mine: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
        """All-gather the branch predictions inside the CFG pair; returns (cond, uncond)."""
        if self.cfg_size == 1:
            raise RuntimeError("exchange_cfg needs cfg_size == 2")
        both = _all_gather_stacked(mine, 2, self.cfg_group)
        return both[0], both[1]

    # ------------------------------------------------------------------------------------ token shards
    def shard_bounds(self, tokens: int) -> Tuple[int, int]:
        if tokens % self.sp_size:
            raise ValueError(f"token count {tokens} is not divisible by the sequence-parallel size {self.sp_size}")
        n = tokens // self.sp_size
        return self.sp_rank * n, (self.sp_rank + 1) * n

    def tokens_to_heads(self, qkv_local: torch.Tensor, heads: int, head_dim: int) -> torch.Tensor:
        """[S/P, 3*H*Dh] (q|k|v column blocks of my tokens) -> [3, S, (H/P)*Dh] (all tokens, my heads).

        Pack so that destination r's data is contiguous -- [P, 3, S/P, (H/P)*Dh] -- then one all_to_all_single;
        the receive buffer [P(src), 3, S/P, hp*Dh] is permuted to [3, P*S/P = S, hp*Dh] (source-rank order is
        token order because shards are contiguous)."""
        P = self.sp_size
        n_local = qkv_local.shape[0]
        hp = heads // P
        if heads % P:
            raise ValueError(f"{heads} heads are not divisible by the sequence-parallel size {P}")
        send = qkv_local.view(n_local, 3, P, hp * head_dim).permute(2, 1, 0, 3).contiguous()
        recv = torch.empty_like(send)
        all_to_all_single(recv, send, self.sp_group)
        return recv.permute(1, 0, 2, 3).reshape(3, P * n_local, hp * head_dim)

    def heads_to_tokens(self, o_heads: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """[S, (H/P)*Dh] (all tokens, my heads) -> [S/P, H*Dh] (my tokens, all heads)."""
        P = self.sp_size
        S, w = o_heads.shape
        n_local = S // P
        send = o_heads.contiguous()                      # chunk r = tokens of rank r, already contiguous
        recv = torch.empty_like(send)                    # [P(src = head group), S/P, hp*Dh]
        all_to_all_single(recv, send, self.sp_group)
        res = recv.view(P, n_local, w).permute(1, 0, 2).reshape(n_local, P * w)
        if out is not None:
            out.copy_(res)
            return out
        return res.contiguous()

    def gather_tokens(self, local: torch.Tensor) -> torch.Tensor:
        """[S/P, C] -> [S, C] on every rank of the sp group (used once per forward for the output head)."""
        if self.sp_size == 1:
            return local
        local = local.contiguous()
        full = torch.empty((self.sp_size * local.shape[0],) + tuple(local.shape[1:]), dtype=local.dtype,
                           device=local.device)
        dist.all_gather_into_tensor(full, local, group=self.sp_group)
        return full

    # ------------------------------------------------------------------------------------ dual-stream (MMDiT) families
    def joint_tokens_to_heads(self, qkv_local: torch.Tensor, img_rows: slice, txt_rows: slice, heads: int, head_dim: int,
                              img_first: bool) -> torch.Tensor:
        """Dual-stream blocks (mmdit.dual_stream_block): the IMAGE / latent stream is token-sharded over the sp group, the
        short text stream is replicated on every rank.  ``qkv_local`` [n_img_local + n_txt, 3*H*Dh] holds my image rows
        and ALL text rows.  Returns [3, S_img + n_txt, (H/P)*Dh] -- all image tokens (Ulysses all-to-all) and all text
        tokens (a local column slice: the text q|k|v of my heads were computed here) of MY heads, concatenated in the
        reference's order (image first for HunyuanVideo-1.5, text first for Flux / QwenImage)."""
        P = self.sp_size
        hp = heads // P
        if heads % P:
            raise ValueError(f"{heads} heads are not divisible by the sequence-parallel size {P}")
        w = hp * head_dim
        img = self.tokens_to_heads(qkv_local[img_rows], heads, head_dim)                  # [3, S_img, w]
        n_txt = qkv_local[txt_rows].shape[0]
        txt = qkv_local[txt_rows].reshape(n_txt, 3, heads, head_dim)[:, :, self.sp_rank * hp:(self.sp_rank + 1) * hp]
        txt = txt.reshape(n_txt, 3, w).permute(1, 0, 2)                                   # [3, n_txt, w]
        return torch.cat([img, txt] if img_first else [txt, img], dim=1).contiguous()

    def joint_heads_to_tokens(self, o_heads: torch.Tensor, n_img_total: int, img_first: bool, out: torch.Tensor,
                              img_rows: slice, txt_rows: slice) -> None:
        """Inverse of ``joint_tokens_to_heads`` for the attention output [S_img + n_txt, (H/P)*Dh]: image rows go back to
        their token owners (all-to-all) into ``out[img_rows]``; the text rows of all head groups are all-gathered so that
        every rank continues the replicated text stream with the full ``out[txt_rows]`` [n_txt, H*Dh]."""
        P = self.sp_size
        S, w = o_heads.shape
        n_txt = S - n_img_total
        img = o_heads[:n_img_total] if img_first else o_heads[n_txt:]
        txt = o_heads[n_img_total:] if img_first else o_heads[:n_txt]
        self.heads_to_tokens(img, out=out[img_rows])
        txt = txt.contiguous()
        allt = torch.empty((P * n_txt, w), dtype=txt.dtype, device=txt.device)      # concatenated along dim 0 (gloo-compatible)
        dist.all_gather_into_tensor(allt, txt, group=self.sp_group)
        out[txt_rows].copy_(allt.view(P, n_txt, w).permute(1, 0, 2).reshape(n_txt, P * w))

    # ------------------------------------------------------------------------------------ frames
    def allgather_frames(self, mine: torch.Tensor) -> torch.Tensor:
        """Single all-gather over ALL ranks of equally shaped per-rank tile stacks -> [world, ...]."""
        if self.world_size == 1:
            return mine.unsqueeze(0)
        return _all_gather_stacked(mine, self.world_size, self.world_group)


class PeerExchange:
    """NVLink peer-mapped (symmetric) buffers for the Ulysses exchange FUSED into the kernels: the q/k RMS-norm+RoPE
    kernel and a scatter copy of v store straight into the receive planes ``qkv[3, S, (H/P)*Dh]`` of the rank that
    owns each head group (``b200_rmsnorm_rope_scatter``), and the attention epilogue stores each output row into the
    ``o[S/P, H*Dh]`` buffer of the rank that owns the token (``b200_attn_fwd_scatter``).  No packing pass, no NCCL
    all-to-all; two device-side barriers per layer order the stores against their consumers.

    Built on ``torch.distributed._symmetric_memory`` (CUDA peer mappings + signal-pad barriers)."""

    def __init__(self, par: "ParallelContext", tokens_total: int, heads: int, head_dim: int, device):
        import ctypes

        import torch.distributed._symmetric_memory as symm_mem

        P = par.sp_size
        if heads % P or tokens_total % P:
            raise ValueError("heads and tokens must be divisible by the sequence-parallel size")
        self.P, self.width, self.n_local = P, (heads // P) * head_dim, tokens_total // P
        self.tokens_total, self.plane = tokens_total, tokens_total * (heads // P) * head_dim
        self.head_off = par.sp_rank * (heads // P)
        self.row0 = par.sp_rank * self.n_local
        self.qkv = symm_mem.empty((3, tokens_total, self.width), dtype=torch.bfloat16, device=device)
        self.o = symm_mem.empty((self.n_local, heads * head_dim), dtype=torch.bfloat16, device=device)
        self.h_qkv = symm_mem.rendezvous(self.qkv, par.sp_group)
        self.h_o = symm_mem.rendezvous(self.o, par.sp_group)
        self.qkv_peers = (ctypes.c_void_p * P)(*[int(p) for p in self.h_qkv.buffer_ptrs])
        self.o_peers = (ctypes.c_void_p * P)(*[int(p) for p in self.h_o.buffer_ptrs])

    def barrier(self, channel: int) -> None:
        """Device-side barrier over the sp group on the current stream (all earlier peer stores are visible after it)."""
        self.h_qkv.barrier(channel=channel)


class JointPeerExchange:
    """PeerExchange for the JOINT sequences of the dual-stream families (``mmdit.dual_stream_block``): the image / latent
    stream is token-sharded, the short text stream is replicated on every rank.

    * ``qkv`` [3, S_img + n_txt, (H/P)*Dh] -- receive planes of MY head group over the whole joint sequence, in the
      reference's concatenation order: every rank stores the q / k / v rows of its image shard straight into the owners'
      planes over NVLink (``b200_rmsnorm_rope_scatter`` in its pure scatter form, after the in-place per-head norm +
      RoPE), the text rows are a local column slice (the text q|k|v of all heads were computed here).
    * ``o`` [n_txt + S_img/P, H*Dh] in MY local joint order -- the attention epilogue of every rank stores the image rows
      of its head group to the token owner and the text rows to EVERY rank (``b200_attn_fwd_scatter_joint``), so the
      replicated text stream continues without the NCCL all-gather of the baseline path.
    Replaces, per block: pack copy + ``all_to_all_single`` + permute + ``cat`` (tokens -> heads) and ``all_to_all_single`` +
    ``all_gather`` + copy (heads -> tokens).  Two device-side barriers per block, as in the Wan path."""

    def __init__(self, par: "ParallelContext", n_img_total: int, n_txt: int, heads: int, head_dim: int, img_first: bool,
                 device):
        import ctypes

        import torch.distributed._symmetric_memory as symm_mem

        P = par.sp_size
        if heads % P or n_img_total % P:
            raise ValueError("heads and image tokens must be divisible by the sequence-parallel size")
        self.P, self.width, self.n_local, self.n_txt = P, (heads // P) * head_dim, n_img_total // P, n_txt
        self.n_img_total, self.img_first, self.S = n_img_total, img_first, n_img_total + n_txt
        self.plane = self.S * self.width
        self.head_off = par.sp_rank * (heads // P)
        self.group_col0 = par.sp_rank * self.width        # my head group's first column inside a [*, H*Dh] block
        self.row0 = par.sp_rank * self.n_local            # my shard's first image row
        self.img_off = 0 if img_first else n_txt          # first image row inside the joint sequence
        self.txt_off = n_img_total if img_first else 0
        self.qkv = symm_mem.empty((3, self.S, self.width), dtype=torch.bfloat16, device=device)
        self.o = symm_mem.empty((self.n_local + n_txt, heads * head_dim), dtype=torch.bfloat16, device=device)
        self.h_qkv = symm_mem.rendezvous(self.qkv, par.sp_group)
        self.h_o = symm_mem.rendezvous(self.o, par.sp_group)
        self.qkv_peers = (ctypes.c_void_p * P)(*[int(p) for p in self.h_qkv.buffer_ptrs])
        self.o_peers = (ctypes.c_void_p * P)(*[int(p) for p in self.h_o.buffer_ptrs])

    def barrier(self, channel: int) -> None:
        """Device-side barrier over the sp group on the current stream (all earlier peer stores are visible after it)."""
        self.h_qkv.barrier(channel=channel)


def _all_gather_stacked(mine: torch.Tensor, n: int, group) -> torch.Tensor:
    """all_gather_into_tensor through flat 1-D views (the form every backend accepts) -> [n, *mine.shape]."""
    flat = mine.contiguous().view(-1)
    out = torch.empty(n * flat.numel(), dtype=mine.dtype, device=mine.device)
    dist.all_gather_into_tensor(out, flat, group=group)
    return out.view((n,) + tuple(mine.shape))


def plan_layout(world_size: int, use_cfg: bool) -> Tuple[int, int]:
    """(cfg_size, sp_size): the CFG pair takes the first factor of 2 when guidance is on."""
    if world_size < 1:
        raise ValueError("world_size must be >= 1")
    cfg = 2 if (use_cfg and world_size % 2 == 0) else 1
    return cfg, world_size // cfg


def all_to_all_single(recv: torch.Tensor, send: torch.Tensor, group) -> None:
    """dist.all_to_all_single with an all-gather based emulation for backends without it (gloo, CPU tests)."""
    backend = dist.get_backend(group)
    if backend == "nccl":
        dist.all_to_all_single(recv, send, group=group)
        return
    P = dist.get_world_size(group)
    me = dist.get_rank(group)
    parts = [torch.empty_like(send) for _ in range(P)]
    dist.all_gather(parts, send.contiguous(), group=group)
    chunk = send.shape[0] // P
    for src in range(P):
        recv[src * chunk:(src + 1) * chunk].copy_(parts[src][me * chunk:(me + 1) * chunk])


def deal_lpt(costs: Sequence[float], world_size: int) -> List[List[int]]:
    """Longest-processing-time-first dealing of items (VAE tiles, cost = latent pixels of the tile) to ranks: items sorted by
    decreasing cost (ties: by index) go to the currently least-loaded rank (ties: lowest rank).  Deterministic, identical on every
    rank.  For the 720p Wan latent (28 tiles: 18 full 32 x 32, 10 smaller edge tiles) on 8 ranks the heaviest rank carries 3 full
    tiles = 7.7 x less than the whole decode, where round-robin gives one rank 4 full tiles (5.8 x).  Returns the sorted item
    lists per rank."""
    order = sorted(range(len(costs)), key=lambda i: (-float(costs[i]), i))
    load = [0.0] * world_size
    mine: List[List[int]] = [[] for _ in range(world_size)]
    for i in order:
        r = min(range(world_size), key=lambda k: (load[k], k))
        mine[r].append(i)
        load[r] += float(costs[i])
    return [sorted(m) for m in mine]


def deal_round_robin(n_items: int, world_size: int, rank: int) -> List[int]:
    """Indices of the items (VAE tiles) rank ``rank`` owns: i with i % world == rank."""
    return [i for i in range(n_items) if i % world_size == rank]
