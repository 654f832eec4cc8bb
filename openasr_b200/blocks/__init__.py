"""Mirror of the reference's ``blocks`` package for the one module on the hot path."""
