from torchdistx_b200.fake import fake_mode, is_fake, meta_like  # noqa: F401

__all__ = ["fake_mode", "is_fake", "meta_like"]
