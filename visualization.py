"""Import-compatibility shim for the reference's ``visualization.py``.

Plotting is outside the hot path (OUT OF SCOPE) and the
reference's module needs matplotlib / seaborn / plotly / LaTeX.  ``Runs.ipynb``
does ``from visualization import *`` (cells 9 and 11) and relies on that star
import for ``np``, so this module exists, exports numpy, and exposes the
reference's entry-point names as stubs that say clearly what they are.
The result dictionaries produced by ``structure.run_experiment`` keep the keys
the real plotting code reads (visualization.py:141-142, :258, :409, :1132, :1249),
so the reference's own visualization.py can be pointed at saved results.
"""
import numpy as np  # noqa: F401  (Runs.ipynb cell 9 uses `np` through the star import)


def _plotting_stub(name):
    def fn(*args, **kwargs):
        raise NotImplementedError(
            f"visualization.{name}: plotting is outside the B200 hot path; use the reference's "
            f"visualization.py on the pickled results (the result-dict keys are unchanged)")
    fn.__name__ = name
    return fn


for _n in ("plot_losses", "plot_metrics_vs_param", "plot_heatmap", "plot_metric_heatmap",
           "plot_metrics_vs_param_grouped", "plot_reconstruction_rows", "plot_gt_accuracy"):
    globals()[_n] = _plotting_stub(_n)
del _n
