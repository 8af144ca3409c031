"""nsa_vibe_b200 -- B200-native (sm_100a) implementation of the NSA hot path of seconds-0/nsa-vibe behind the
reference's own module interface.  See DESIGN.md and INTEGRATION.md."""
from .cache.kv_cache import NSA_KV, create_empty_kv  # noqa: F401
from .core.block_index import BlockMeta, build_block_meta  # noqa: F401
from .core.nsa_attention import GateMLP, NSAAttention  # noqa: F401
from .ops import NSAConfig  # noqa: F401

__all__ = ["NSAAttention", "GateMLP", "NSA_KV", "create_empty_kv", "BlockMeta", "build_block_meta", "NSAConfig"]
