"""B200-native measurement-operator hot path of pfb-imaging (w-stacked gridder /
degridder + Hessian apply), behind the reference's operator call signatures."""

__version__ = "0.1.0"
