"""torchdistx_b200 -- Blackwell-native deferred initialisation (drop-in for torchdistx.fake /
torchdistx.deferred_init).  See DESIGN.md."""
__version__ = "0.1.0"
