"""topsicle_b200 -- B200-native per-read telomere scan (drop-in for Topsicle's hot path)."""
from .patterns import pattern_scramble_telo, patterns_to_search  # noqa: F401

__version__ = "0.1.0"
