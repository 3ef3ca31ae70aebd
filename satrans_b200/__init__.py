"""satrans_b200 — B200-native SATrans training/scoring hot path (drop-in for the reference's
`models/satrans.py` on `models/meta_basemodel.py`)."""
