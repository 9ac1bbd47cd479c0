"""krets_b200 — B200-native batched solver backend for the krets circuit simulator.

Python here is a thin ctypes veneer over the C ABI (include/krets_b200.h) and the C++ host mirror of the krets API
(include/krets_b200_host.h).  All numerics run in hand-written sm_100a CUDA kernels; there is no CPU fallback.
"""
from .api import (DEVICES_LEVEL1, DEVICES_REFERENCE, AcAnalysis, AcSweep, Analysis, AnalysisResult, Circuit, DcAnalysis, KretsError, Plan, Solver,
                  SolverConfig, TransientAnalysis, device_count, generate_frequencies, kernel_launch_count,
                  last_kernel_ms, parse_circuit_description, parse_circuit_description_file, version)

__all__ = ["DEVICES_LEVEL1", "DEVICES_REFERENCE", "AcAnalysis", "AcSweep", "Analysis", "AnalysisResult", "Circuit", "DcAnalysis", "KretsError", "Plan", "Solver",
           "SolverConfig", "TransientAnalysis", "device_count", "generate_frequencies", "kernel_launch_count",
           "last_kernel_ms", "parse_circuit_description", "parse_circuit_description_file", "version"]
