"""depthmodelhardening_b200 -- B200-native hot path for DepthModelHardening."""
__version__ = "0.1.0"
