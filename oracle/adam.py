"""Oracle (test infrastructure): tf.train.AdamOptimizer update (models.py:168,178).

TF1 semantics (beta1 0.9, beta2 0.999, epsilon 1e-8): the bias correction is
folded into the step size and epsilon is added to the UNcorrected sqrt(v)
("epsilon hat"), which differs from torch.optim.Adam:

    lr_t = lr * sqrt(1 - beta2^t) / (1 - beta1^t)
    m <- beta1 m + (1 - beta1) g ;  v <- beta2 v + (1 - beta2) g^2
    theta <- theta - lr_t * m / (sqrt(v) + eps)
"""
import numpy as np


def adam_tf_step(theta, g, m, v, step, lr=1e-3, b1=0.9, b2=0.999, eps=1e-8):
    """step is 1-based.  Returns (theta, m, v) new arrays, dtype of theta."""
    dt = theta.dtype
    lr_t = lr * np.sqrt(1.0 - b2 ** step) / (1.0 - b1 ** step)
    m = (b1 * m + (1.0 - b1) * g).astype(dt)
    v = (b2 * v + (1.0 - b2) * g * g).astype(dt)
    theta = (theta - lr_t * m / (np.sqrt(v) + eps)).astype(dt)
    return theta, m, v
