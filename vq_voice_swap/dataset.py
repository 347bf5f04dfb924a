from vq_voice_swap_b200.audio_io import (ChunkReader, ChunkWriter, decode_to_linear, decode_u_law,  # noqa: F401
                                         encode_from_linear, encode_u_law)
