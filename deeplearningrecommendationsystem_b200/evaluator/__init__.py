"""Mirror of the reference evaluator package (evaluator/evaluator.py, evaluator/ranking.py)."""
from .evaluator import Evaluator, binary_metrics  # noqa: F401
from .ranking import Ranking  # noqa: F401
