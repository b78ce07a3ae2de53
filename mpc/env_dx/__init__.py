"""Drop-in import path of the reference's example systems (mpc/env_dx); see mpc/pytorch_b200/dynamics.py."""
