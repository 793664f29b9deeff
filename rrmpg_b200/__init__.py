"""rrmpg_b200 -- B200-native ensemble rainfall-runoff engine (drop-in for rrmpg.models hot path)."""
