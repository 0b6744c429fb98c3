"""nans_projekat_b200 — B200-native rigid-body step behind the reference's nans.so plugin API."""
