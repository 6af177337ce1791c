"""oprl_b200 -- B200-native off-policy update engine behind oprl's Algorithm / ReplayBuffer API."""
__version__ = "0.1.0"
