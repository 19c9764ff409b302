# coding: utf-8
"""plain/ variant of the reference (isotropic Gaussian latent instead of the ABCD sampler; /root/reference plain/):
same encoder / decoder kernels, `model.Sampler` as the latent (two MLPs -> mean, log-variance; fused reparameterised
sample + KL kernel).  Entry points: plain.learning (plain/learning.py) and plain.encode (plain/encode.py)."""
