"""seaduck_b200 -- B200-native implementation of seaduck's Lagrangian / interpolation hot path."""
