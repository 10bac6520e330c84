"""B200-native (sm_100a) post-network decoder of OffsetGuided: ``offsetguided_b200.decoder`` mirrors
the reference's ``decoder`` package, ``engine.DecoderEngine`` owns one handle of the C ABI
(``include/og_decoder.h``, ``libogdecoder.so``).  Build with ``python -m offsetguided_b200.build``."""
