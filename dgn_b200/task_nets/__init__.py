"""Task networks (callers of the hot path) on the fused DGN layers.

They live outside the ``nets`` namespace on purpose: with the overlay of INTEGRATION.md the reference's
own ``nets/<task>/dgn_net.py`` files must keep resolving to the reference.
"""
