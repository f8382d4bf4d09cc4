nm = None
