"""zs3.utils: loss.py and metrics.py are the B200-native ones; everything else (saver, lr_scheduler, summaries,
calculate_weights: outside the hot path) falls through to the reference checkout when it is on sys.path."""
from zs3 import _extend_search_path

_extend_search_path(__path__, ("zs3", "utils"))
