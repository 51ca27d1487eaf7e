"""CPU module placeholder (``_libepseon_cpu``: ``greet`` only, as in the reference)."""
