"""farkle_ii_b200 — B200-native engine for the Farkle_II simulation hot path.

Host-side mirror of the reference interface (``farkle.simulation.run_tournament``,
``farkle.simulation.simulation``, ``farkle.simulation.strategies``,
``farkle.utils.random``, the H2H ``BlockRunner``) over the C ABI declared in
``include/farkle_b200.h``.  Importing the package never touches CUDA; the first
compute call binds a device and raises if none is available.
"""

__version__ = "0.1.0"
