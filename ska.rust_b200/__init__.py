"""ska.rust_b200 -- B200-native hot path of bacpop/ska.rust (split-k-mer build / merge / filter / distance).

The directory name contains a dot, so import it through the top-level shim:  `import ska_rust_b200`.
Layout:  csrc/ (CUDA kernels + the C ABI of include/skagpu.h -> libskagpu.so), api.py (host-side mirror
of the reference interface over ctypes), host/ (C++ host: .skf codec, FASTA/FASTQ reader, `ska` CLI).
"""
from .api import *  # noqa: F401,F403
from .api import SkaError, MergeSkaArray, QualOpts, Sample, SeqBatch, build_and_merge, build_from_samples  # noqa: F401
