"""
Data ingest for the drug-sensitivity matrices (mirror of the reference's
data/drug_sensitivity/load_data.py:37-68): tab-separated text, rows = cell lines,
columns = drugs, unobserved entries written as nan.  The nan -> (R=0, M=0)
conversion is one vectorised pass (the reference walks all I*J cells in Python,
load_data.py:41-44).  The datasets themselves are not shipped with this package:
every loader takes the file location.
"""
import numpy

DELIM = '\t'
MIN_GDSC = 3  # minimum number of observed drugs per cell line


def load_data_create_mask(location):
    ''' Load in .txt file, and set mask entries for nan to 0. '''
    R = numpy.loadtxt(location, dtype=float, delimiter=DELIM, ndmin=2)
    missing = numpy.isnan(R)
    M = numpy.where(missing, 0., 1.)
    R[missing] = 0.
    return (R, M)


def load_gdsc_ic50(location):
    ''' Return (R_gdsc, M_gdsc). Filter out cell lines with fewer than :MIN_GDSC observed entries. '''
    R, M = load_data_create_mask(location)
    keep = M.sum(axis=1) >= MIN_GDSC
    return R[keep, :], M[keep, :]


def load_ctrp_ec50(location):
    ''' Return (R_ctrp, M_ctrp). '''
    return load_data_create_mask(location)


def load_ccle_ic50(location):
    ''' Return (R_ccle_ic50, M_ccle_ic50). '''
    return load_data_create_mask(location)


def load_ccle_ec50(location):
    ''' Return (R_ccle_ec50, M_ccle_ec50). '''
    return load_data_create_mask(location)
