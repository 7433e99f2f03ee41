"""Mirror of the reference's `acmpc.utils` for the modules on the MPC path (load)."""
