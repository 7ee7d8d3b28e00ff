"""CPU: the oracle against cv2 known answers at parameter extremes and awkward shapes (tests/golden/kat_extreme.npz)."""
import extreme_cases as X


def test_oracle_lk_extreme(oracle, golden):
    X.check_lk_extreme(oracle, golden("kat_extreme.npz"))


def test_oracle_gftt_extreme(oracle, golden):
    X.check_gftt_extreme(oracle, golden("kat_extreme.npz"))


def test_oracle_tiny(oracle, golden):
    X.check_tiny(oracle, golden("kat_extreme.npz"))
