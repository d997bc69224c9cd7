"""mop_b200 - B200 (sm_100a) implementation of the MoP attention hot path.

Drop-in ``nn.Module`` replacements for the attention classes of Eran-BA/MoP,
backed by hand-written CUDA kernels behind a C ABI (``libmop_b200.so``,
``include/mop_b200.h``).  There is no CPU fallback.
"""
from .attention_variants import BaselineMSA, CrossViewMixerMSA, EdgewiseGateHead, EdgewiseMSA, MultiHopMSA, UnifiedMSA
from .components import MLP, MSA, Block, DropPath, PatchEmbed
from .functional import edgewise_attention, quartet_attention, sdpa
from .quartet_attn_patch import CausalSelfAttention, TransformerConfig
from .vit_edgewise import BlockEdgewise, ViTEdgewise
from .whisper_mop import MoP2D, MultiheadCrossAttention, MultiheadSelfAttention

__all__ = [
    "BaselineMSA", "CrossViewMixerMSA", "MultiHopMSA", "EdgewiseGateHead", "EdgewiseMSA", "UnifiedMSA", "MSA", "MLP", "Block", "DropPath", "PatchEmbed",
    "BlockEdgewise", "ViTEdgewise", "MultiheadSelfAttention", "MultiheadCrossAttention", "MoP2D",
    "CausalSelfAttention", "TransformerConfig", "edgewise_attention", "quartet_attention", "sdpa",
]
