"""Drop-in `models` package: same module names, constructors, forward signatures and state_dict
keys as cansdev/wireframe-3d-prediction's `models/`, computed by libwf_b200.so on a B200."""
