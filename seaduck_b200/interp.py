"""Host driver of the device interpolation kernel (K3); see csrc/interp.cu."""


def interpolate(pos, var_name, knw, vec_transform=True, prefetched=None, prefetch_prefix=None):
    raise NotImplementedError("general kernel interpolation (sdk_interp) is not built yet")
