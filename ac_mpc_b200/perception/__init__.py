"""Mirror of the reference's `acmpc.perception` for the functions that feed the MPC step (track smoothing)."""
