"""expertsim (B200-native): drop-in for the hot path of the reference ExpertSim package — same module / function names
(`expertsim.models`, `expertsim.train`, `expertsim.config`), CUDA kernels underneath (see include/expertsim_b200.h)."""
__version__ = "0.1.0"
