"""Gradient clipping, tower averaging and TF-1.x optimizer update rules.
TEST INFRASTRUCTURE.

Follows ``models/model_base.py:12-20`` (optimizer table), ``:68-95``
(``_set_optimizer``: momentum/nesterov use 0.9, others TF defaults),
``:135-166`` (``_clip_gradients`` = per-tensor ``tf.clip_by_norm``) and
``utils/training/multi_gpu.py:13-48`` (``average_gradients`` = unweighted mean
over towers).  Update rules are TF-upstream, not vendored:

  sgd       : w -= lr*g
  momentum  : a = 0.9*a + g ;            w -= lr*a
  nestrov   : a = 0.9*a + g ;            w -= lr*(g + 0.9*a)
  adagrad   : a(0)=0.1 ; a += g^2 ;      w -= lr*g/sqrt(a)
  adadelta  : rho=.95 eps=1e-8 ; a = rho*a+(1-rho)g^2 ; u = sqrt(d+eps)/sqrt(a+eps)*g ;
              d = rho*d+(1-rho)u^2 ;     w -= lr*u
  adam      : b1=.9 b2=.999 eps=1e-8 ; lr_t = lr*sqrt(1-b2^t)/(1-b1^t) ;
              m = b1*m+(1-b1)g ; v = b2*v+(1-b2)g^2 ; w -= lr_t*m/(sqrt(v)+eps)
  rmsprop   : decay=.9 momentum=0 eps=1e-10 ; ms(0)=1 ; ms = .9*ms+.1*g^2 ;
              mom = 0*mom + lr*g/sqrt(ms+eps) ; w -= mom
"""
import numpy as np

OPTIMIZERS = ("adagrad", "adadelta", "adam", "rmsprop", "sgd", "momentum", "nestrov")


def clip_by_norm(g, clip_norm):
    """tf.clip_by_norm: g * clip / max(||g||_2, clip)."""
    n = np.sqrt(np.sum(np.square(g.astype(np.float64))))
    return (g * (clip_norm / max(n, clip_norm))).astype(g.dtype)


def average_gradients(tower_grads):
    """tower_grads: list (towers) of lists (variables) of arrays (or None).
    Mean over the towers that have a gradient (multi_gpu.py:30-40)."""
    out = []
    for per_var in zip(*tower_grads):
        gs = [g for g in per_var if g is not None]
        out.append(np.mean(np.stack(gs, 0), axis=0))
    return out


class Optimizer(object):
    def __init__(self, name, learning_rate):
        name = name.lower()
        if name not in OPTIMIZERS:
            raise ValueError("Optimizer name should be one of [%s], you provided %s." %
                             (", ".join(OPTIMIZERS), name))
        self.name, self.lr, self.t, self.state = name, learning_rate, 0, {}

    def step(self, params, grads, learning_rate=None):
        """In-place update of ``params`` (list of float arrays)."""
        lr = self.lr if learning_rate is None else learning_rate
        self.t += 1
        n = self.name
        for k, (w, g) in enumerate(zip(params, grads)):
            if g is None:
                continue
            st = self.state.setdefault(k, {})
            if n == "sgd":
                w -= lr * g
            elif n in ("momentum", "nestrov"):
                a = st.setdefault("a", np.zeros_like(w))
                a[...] = 0.9 * a + g
                w -= lr * (g + 0.9 * a) if n == "nestrov" else lr * a
            elif n == "adagrad":
                a = st.setdefault("a", np.full_like(w, 0.1))
                a += g * g
                w -= lr * g / np.sqrt(a)
            elif n == "adadelta":
                a = st.setdefault("a", np.zeros_like(w))
                d = st.setdefault("d", np.zeros_like(w))
                a[...] = 0.95 * a + 0.05 * g * g
                u = np.sqrt(d + 1e-8) / np.sqrt(a + 1e-8) * g
                d[...] = 0.95 * d + 0.05 * u * u
                w -= lr * u
            elif n == "adam":
                m = st.setdefault("m", np.zeros_like(w))
                v = st.setdefault("v", np.zeros_like(w))
                lr_t = lr * np.sqrt(1 - 0.999 ** self.t) / (1 - 0.9 ** self.t)
                m[...] = 0.9 * m + 0.1 * g
                v[...] = 0.999 * v + 0.001 * g * g
                w -= lr_t * m / (np.sqrt(v) + 1e-8)
            elif n == "rmsprop":
                ms = st.setdefault("ms", np.ones_like(w))
                ms[...] = 0.9 * ms + 0.1 * g * g
                w -= lr * g / np.sqrt(ms + 1e-10)
