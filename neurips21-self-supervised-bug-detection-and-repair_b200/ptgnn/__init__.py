"""B200-native implementation of the ``ptgnn`` operator surface that BugLab's gnn-mlp path imports.

``ptgnn`` is an unpinned third-party dependency of the reference (requirements.txt:13) that holds the
arithmetic of the hot path; it is absent from /root/reference and not installable offline.  This package
keeps its import paths and class names (checkpoints are pickles of these classes — reference
buglab/data/modelsync/server.py:34, buglab/models/modelregistry.py:154) and re-implements them B200-first:
host code prepares packed int32 tables in pinned memory, every arithmetic step runs in the sm_100a kernels
behind ``include/buglab_b200.h``.  There is no CPU compute path.
"""
