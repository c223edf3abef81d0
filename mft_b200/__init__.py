"""mft_b200: Blackwell (sm_100a) implementation of the MFT per-frame tracking hot path.

Python surface mirrors the reference (serycjon/MFT): ``mft_b200.MFT.MFT`` (init/track),
``mft_b200.raft.RAFTWrapper`` (compute_flow), ``mft_b200.results.FlowOUTrackingResult``,
``mft_b200.config.Config``.  All arithmetic runs in hand-written CUDA behind the C ABI of
``libmft_b200.so`` (include/mft_b200.h); there is no CPU fallback."""
__version__ = '0.1'
