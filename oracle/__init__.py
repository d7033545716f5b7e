"""CPU oracle for the ConvLSTM recurrence hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the product:
only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline``
/ ``--impl reference`` legs may import it, and only as the checker / the timed
CPU baseline.  The product package (``pl-convlstm-gan_b200``) never imports it
and fails loudly when its CUDA library is missing.

Parity status: the reference repo ships no golden vectors for this path
("parity unpinned" by the reference's own tests), so the
oracle is pinned against outputs of the reference itself, generated in the build
container by ``tests/golden/make_golden.py`` (imports ``/root/reference``
unmodified) and committed under ``tests/golden/*.npz``.
"""
from .convlstm_oracle import (  # noqa: F401
    cell_forward,
    cell_forward_gates,
    cell_backward,
    stack_forward,
    stack_backward,
    encoder_forecaster_forward,
    conv2d_same,
    add_coord_channels,
    frontend_forward,
    nowcast_forward,
)
from . import loss_oracle  # noqa: F401,E402
