"""Modality constants (reference: data/schema.py:3-13). MODALITIES order == mask-column order everywhere."""
from typing import Dict, List

MODALITIES: List[str] = ["clinical", "datspect", "mri"]
MODALITY_FEATURES: Dict[str, List[str]] = {
    "clinical": ["age", "sex", "education", "updrs_iii", "disease_duration"],
    "datspect": ["caudate_l", "caudate_r", "putamen_l", "putamen_r", "sbr_mean"],
    "mri": ["hippocampus_l", "hippocampus_r"],
}
TARGET_COL = "diagnosis"
ID_COL = "patno"
