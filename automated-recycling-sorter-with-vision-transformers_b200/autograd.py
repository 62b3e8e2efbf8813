"""Training path (forward that saves activations + hand-written backward). Filled in below."""
from ._lib import VitkError


def encoder_forward_train(module, images):
    raise VitkError("training path is not built yet - wrap inference in torch.no_grad()")


def classifier_forward_train(module, images):
    raise VitkError("training path is not built yet - wrap inference in torch.no_grad()")
