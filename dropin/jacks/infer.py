from jacks_b200.infer import LOG, inferJACKS, inferJACKSGene, infer_batched  # noqa: F401
