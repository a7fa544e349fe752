"""CLI of the drop-in: ``python main.py <QWS|qws|Normal> <ML|PNLow|PNHigh|ML+2PN|WOA|ESWOA> [epoch ...]``.

Same dispatch contract as the reference's ``main.py:14-84,216-228`` for the four modes on the accelerated
path: the ``environment.ini`` section ``<dataset>-<approach>`` is read and its values are passed
POSITIONALLY (the ini key order is API), trailing epoch selectors can be overridden from argv, and PNHigh
receives the ``epochML`` value for both ``epochML`` and ``epochPNLow`` (reference quirk, main.py:68,83).
``WOA`` (= ML+2PN+WOA: ESWOA fine-tuning seeded by the PNHigh picks, main.py:86-104) and ``ESWOA`` (main.py:126-140)
run the search with its fitness evaluations on the GPU; the other CPU metaheuristic baselines (ML+ESWOA, DAAGA, SDFGA,
DPKSD, PDDQN) are outside this repository's scope.
"""
import configparser
import sys

import src.ML2PN as ML2PN
import src.baselines.WOA as WOA
import src.models.trainML as trainML
import src.models.trainPNHigh as trainPNHigh
import src.models.trainPNLow as trainPNLow

INT, FLT = int, float
# positional casts per approach (main.py:23-24, 39-42, 65-68, 221)
CASTS = {
    "ML": [INT, INT, INT, INT, FLT, FLT, INT],
    "PNLow": [INT] * 9 + [FLT, FLT, FLT, INT],
    "PNHigh": [INT] * 9 + [FLT, FLT, FLT, INT, INT],
    "ML+2PN": [INT, INT],
    "WOA": [INT] * 6 + [FLT, INT, INT, INT],          # reduct is a float in the Normal sections (main.py:102)
    "ESWOA": [INT] * 6 + [FLT, INT, INT, INT],
}
OUT_OF_SCOPE = {"DAAGA", "SDFGA", "DPKSD", "PDDQN", "ML+ESWOA", "ML+DAAGA", "ML+SDFGA", "ML+DPKSD", "ML+PDDQN"}


def section_values(config, dataset, approach, argv):
    name = f"{dataset}-{approach}"
    values = [config.get(name, key) for key in config.options(name)]
    extra = argv[3:]
    if approach == "PNLow" and len(extra) == 1:
        values[-1] = extra[0]
    elif approach == "PNHigh":
        if len(extra) >= 1:
            values[-1] = extra[0]
        if len(extra) == 2 and int(values[-1]) != -1:
            values[-2] = extra[1]
    elif approach == "ML+2PN" and len(extra) == 1:
        values[-1] = extra[0]
    elif approach == "WOA" and len(extra) == 1:
        values[-3] = extra[0]                                        # epoch (main.py:89-90)
    return [cast(v) for cast, v in zip(CASTS[approach], values)]


def main(argv):
    if len(argv) < 3:
        raise SystemExit(__doc__)
    dataset = "QWS" if argv[1] in ("QWS", "qws") else argv[1]
    approach = argv[2]
    if approach in OUT_OF_SCOPE:
        raise SystemExit(f"{approach}: CPU baseline of the reference, not part of the B200 hot path (see DESIGN.md)")
    if dataset not in ("QWS", "Normal") or approach not in CASTS:
        print("Please check the parameters!")                        # main.py:230-231
        return
    config = configparser.RawConfigParser()
    config.read("environment.ini")
    v = section_values(config, dataset, approach, argv)
    if approach == "ML":
        trainML.TrainML(dataset, *v).start()
    elif approach == "PNLow":
        trainPNLow.PNLow(dataset, *v).start()
    elif approach == "PNHigh":
        trainPNHigh.PNHigh(dataset, *v[:12], v[12], v[12]).start()   # epochML passed twice (main.py:68)
    elif approach in ("WOA", "ESWOA"):
        WOA.WOA(dataset, *v).start()
    else:
        ML2PN.check(dataset, v[0], v[1])


if __name__ == "__main__":
    main(sys.argv)
