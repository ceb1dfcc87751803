"""divortio-lz4_b200: B200-native LZ4 block codec behind the divortio-lz4 block/frame interface.

Import as `divortio_lz4_b200` (see the shim package of that name at the repository root).
"""
