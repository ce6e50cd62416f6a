"""Measured prototypes that are not on the product path ("tensor-core DFT")."""
