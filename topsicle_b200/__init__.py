"""topsicle_b200 -- B200-native per-read telomere scan (drop-in for Topsicle's hot path).

`from topsicle_b200 import *` gives what `from Topsicle import *` gives for this path: the reference's package
star-exports `Topsicle.allsteps` (Topsicle/__init__.py:1); its second star-export, the figure module
`Topsicle.descriptive_plot` (:2), is represented by the data half of its heat map (`descriptive.py`)."""
from .allsteps import *  # noqa: F401,F403
from .allsteps import __all__ as _allsteps_all
from .descriptive import patterns_vs_match_heatmap  # noqa: F401

__version__ = "0.2.0"
__all__ = list(_allsteps_all) + ["patterns_vs_match_heatmap"]
