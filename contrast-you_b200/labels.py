"""Host-side label derivation for the InfoNCE hooks: group strings -> integer class ids -> in-kernel positive mask.

Mirrors ``semi_seg/epochers/helper.py:54-71`` (generator classes) and ``semi_seg/hooks/utils.py:20-102`` (dispatch on
dataset name and ``contrast_on``).  Integer work, bit-exact with the reference; sklearn's ``LabelEncoder`` is
replaced by its definition (rank among the sorted distinct values) so the hot path does not import sklearn.
"""
from functools import lru_cache
from typing import List

__all__ = ["PartitionLabelGenerator", "PatientLabelGenerator", "ACDCCycleGenerator", "SIMCLRGenerator",
           "global_label_generator", "get_label"]


def _rank_encode(values: List[str]) -> List[int]:
    order = {v: i for i, v in enumerate(sorted(set(values)))}
    return [order[v] for v in values]


class PartitionLabelGenerator:
    def __call__(self, partition_list: List[str], **kwargs):
        return _rank_encode(list(partition_list))


class PatientLabelGenerator:
    def __call__(self, patient_list: List[str], **kwargs):
        return _rank_encode(list(patient_list))


class ACDCCycleGenerator:
    def __call__(self, experiment_list: List[str], **kwargs):
        return [0 if e == "00" else 1 for e in experiment_list]


class SIMCLRGenerator:
    def __call__(self, partition_list: List[str], **kwargs):
        return list(range(len(partition_list)))


_COMMON = {"partition": PartitionLabelGenerator, "patient": PatientLabelGenerator, "self": SIMCLRGenerator}
_PER_DATASET = {
    "acdc": dict(_COMMON, cycle=ACDCCycleGenerator),
    "prostate": _COMMON, "mmwhs": _COMMON, "spleen": _COMMON, "hippocampus": _COMMON,
}


@lru_cache()
def global_label_generator(dataset_name: str, contrast_on: str):
    if dataset_name == "acdc" or "acdc" in dataset_name:
        table = _PER_DATASET["acdc"]
    elif dataset_name in ("prostate", "prostate_md"):
        table = _PER_DATASET["prostate"]
    elif dataset_name in _PER_DATASET:
        table = _PER_DATASET[dataset_name]
    else:
        raise NotImplementedError(dataset_name)
    if contrast_on not in table:
        raise NotImplementedError(contrast_on)
    return table[contrast_on]()


def get_label(contrast_on, data_name, partition_group, label_group):
    """semi_seg/hooks/utils.py:74-102: ACDC / prostate group names look like "patient007_01" (patient _ experiment)."""
    if data_name == "acdc" or "acdc" in data_name:
        return global_label_generator(dataset_name="acdc", contrast_on=contrast_on)(
            partition_list=partition_group,
            patient_list=[p.split("_")[0] for p in label_group],
            experiment_list=[p.split("_")[1] for p in label_group])
    if data_name in ("prostate", "prostate_md"):
        return global_label_generator(dataset_name="prostate", contrast_on=contrast_on)(
            partition_list=partition_group, patient_list=[p.split("_")[0] for p in label_group])
    if data_name in ("mmwhsct", "mmwhsmr"):
        return global_label_generator(dataset_name="mmwhs", contrast_on=contrast_on)(
            partition_list=partition_group, patient_list=label_group)
    if data_name in ("spleen", "hippocampus"):
        return global_label_generator(dataset_name=data_name, contrast_on=contrast_on)(
            partition_list=partition_group, patient_list=label_group)
    raise NotImplementedError(data_name)
