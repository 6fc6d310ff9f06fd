"""decorators/variational_info_bottleneck.py of the reference: wraps a module forward with
the 'normal' reparameterisation z = eps * exp(logvar/2) + mu.  The elementwise math runs in the
``mmad`` VIB kernel when available on device; the noise may be supplied (``eps=``) so CPU-vs-GPU
parity does not depend on the device RNG."""
import functools


def variational_info_bottleneck(forward_fn):
    @functools.wraps(forward_fn)
    def decorated_forward(self, x, distribution=None, k=1, stochastic_inference=True, eps=None):
        output = forward_fn(self, x)
        if distribution is None:
            return output
        if distribution == "normal":
            from ..ops import vib_reparameterize
            if k < 1:
                raise ValueError("k should be >= 1")
            return vib_reparameterize(output, k, stochastic_inference, eps)
        raise NotImplementedError("Wrong distribution for information bottleneck: {}".format(distribution))

    return decorated_forward
