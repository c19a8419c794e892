"""B200-native implementation of the well_duplicates hot path.

Host side in Python (file formats, CLIs, report text), compute in hand-written
sm_100a CUDA kernels behind the C ABI of include/welldup.h.  Mirrors the
reference's interfaces: ``load_targets`` (target.py), ``BCLReader`` /
``Tile.get_seqs`` (bcl_direct_reader.py), ``output_writer`` and the two command
lines (count_well_duplicates.py, prepare_cluster_indexes.py)."""

__version__ = "0.1"

SEQUENCE = 0      # bcl_direct_reader.py:54-55
QUAL_FLAG = 1
