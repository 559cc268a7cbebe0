from torchdistx_b200.deferred_init import (  # noqa: F401
    deferred_init,
    is_deferred,
    materialize_module,
    materialize_tensor,
)

__all__ = ["deferred_init", "is_deferred", "materialize_module", "materialize_tensor"]
