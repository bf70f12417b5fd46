"""simplenerf_b200: B200-native (sm_100a) implementation of SimpleNeRF's volumetric-rendering hot path,
a drop-in behind the reference's ``src/models`` forward API.  See DESIGN.md / INTEGRATION.md."""
__version__ = '0.1.0'
