"""cesm_emulator_b200 -- B200-native (sm_100a) hot path of the CESM diffusion emulator."""
__version__ = "0.1.0"
