"""Config-dict access with the reference's error behaviour: a missing key is reported with
``logger.error`` and a bare ``ValueError`` (e.g. reference mapping/grid_maps.py:13-24,
sensors/sensor_factories.py:27-47)."""
import logging
from typing import Any, Dict, Sequence

logger = logging.getLogger(__name__)


def require(params: Dict, path: Sequence[str], what: str = None) -> Any:
    """``params[path[0]][path[1]]...`` or log + ValueError naming the first missing key."""
    node = params
    for depth, key in enumerate(path):
        if not isinstance(node, dict) or key not in node:
            where = "config file" if depth == 0 else "'" + ".".join(path[:depth]) + "' section of the config file"
            logger.error(f"Cannot find {what or key!r} specification ({'.'.join(path)}) in {where}!")
            raise ValueError(f"missing config key {'.'.join(path)}")
        node = node[key]
    return node


def require_member(value: str, known: Sequence[str], kind: str) -> None:
    if value not in known:
        logger.error(f"'{value}' not in list of known {kind}: {list(known)}")
        raise ValueError(f"unknown {kind} '{value}'")
