"""Minimal host-side stand-in for the parts of ``dpu_utils`` the BugLab gnn-mlp path imports
(``RichPath``, ``run_and_debug``, ``Vocabulary``, ``split_identifier_into_parts``).  dpu-utils is an
unpinned third-party dependency of the reference (Dockerfile:14) that is not installable offline; this is
a from-scratch restatement of its documented behaviour for local files only (no Azure storage)."""
