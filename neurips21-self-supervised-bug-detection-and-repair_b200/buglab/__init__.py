"""Drop-in mirror of the BugLab modules on the gnn-mlp hot path (reference: microsoft/neurips21-self-supervised-
bug-detection-and-repair, ``buglab/``).  Same import paths, class names, attribute / state_dict names and entry
points (``python -m buglab.models.train``, ``python -m buglab.models.evaluate``) — re-implemented from scratch on
the buglab_b200 kernels.  Data extraction, rewriting and the self-supervised controllers are not part of this
package (SURVEY.md §2 marks them out of scope)."""
