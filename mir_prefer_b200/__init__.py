"""mir_prefer_b200 -- B200-native fold stage for miR-PREFeR (libmirfold.so + Python host layer)."""
from ._lib import MirfoldError, LIB_PATH  # noqa: F401
from .fold import MirFold, FoldResult, parse_rnalfold_input, format_record, convert_sequence  # noqa: F401

__all__ = ["MirFold", "FoldResult", "MirfoldError", "parse_rnalfold_input", "format_record", "convert_sequence"]
