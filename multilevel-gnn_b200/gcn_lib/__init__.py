"""Drop-in mirror of the reference's layer library (models/gcn_lib) on the sm_100a kernels."""
