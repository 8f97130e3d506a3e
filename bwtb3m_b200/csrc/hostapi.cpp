// placeholder: replaced by the file-level host API (formats + b3m_compute_bwt)
