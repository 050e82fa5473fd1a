# -*- coding: utf-8 -*-
"""String helpers shared with the reference (tricolour/util.py:78-95)."""
import re

import numpy as np


def casa_style_range(val, opt_unit="m"):
    """Parses a CASA style range such as ``"0~550"`` or ``"0~550m"``.

    Blank or ``"*"`` means everything: ``(0, inf)``.  Raises ``ValueError`` on
    anything else that is not ``<number>~<number>[m]``.
    """
    if not isinstance(val, str):
        raise ValueError("Value must be a string")
    if val.strip() == "" or val.strip() == "*":
        return (0, np.inf)
    number = r"(\d+(\.\d*)?|\.\d+)([eE][+-]?\d+)?"
    if re.match(r"^" + number + "~" + number + r"[\s]*[" + opt_unit + "]?$", val):
        val = val.replace(" ", "").replace("\t", "").replace(opt_unit, "")
        return list(map(float, val.split("~")))
    raise ValueError("Value must be range or blank")
