"""Run-time switches (reference ``seaduck/runtime_conf.py``).

The reference's only switch is numba on/off (``rcParam["compilable"]``).  Here the compute backend
is always the CUDA library; ``rcParam["backend"]`` documents that and ``compilable`` is kept so that
code written against the reference keeps working.
"""

rcParam = {
    "compilable": True,
    "dump_masks_to_local": False,
    "backend": "cuda-sm_100a",
}


def compileable(func):
    """Kept for source compatibility with the reference (``runtime_conf.py:14``): a no-op."""
    return func


def compileable_parallel(func):
    return func
