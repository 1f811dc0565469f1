"""textocvp_b200 -- B200-native (sm_100a) TextOCVP rollout path."""
