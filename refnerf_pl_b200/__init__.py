"""refnerf_pl_b200: B200-native (sm_100a) implementation of the Ref-NeRF per-ray rendering hot path of
minfenli/refnerf-pl, behind the reference's own Python surface (`models.Model`, `NerfMLP`, `PropMLP`,
`utils.Rays`, gin names).  See DESIGN.md / INTEGRATION.md."""
__all__ = ['configs', 'utils', 'synthetic']
