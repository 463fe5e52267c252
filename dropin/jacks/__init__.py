"""Import shim: put this directory's parent (``dropin/``) and the repository root on PYTHONPATH and existing
code that does ``from jacks.infer import inferJACKS`` / ``import jacks.jacks_io`` runs on jacks_b200 unchanged."""
