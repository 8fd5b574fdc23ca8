"""Mirror of `nr3d_lib.graphics` for the NeuS rendering path: pack_ops, raysample, raytest, raymarch, nerf, neus."""
