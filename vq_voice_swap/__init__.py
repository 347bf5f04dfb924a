"""Drop-in namespace: the reference's scripts (`sample_diffusion.py`, `sample_vqvae.py`) import
`vq_voice_swap.*`; with this repository first on PYTHONPATH those imports resolve to the sm_100a
implementation in `vq_voice_swap_b200` without touching the scripts."""
