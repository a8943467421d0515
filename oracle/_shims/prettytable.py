"""Stand-in for prettytable (imported by utils/utilities.py, used only by count_parameters)."""


class PrettyTable(object):
    def __init__(self, *a, **k):
        self.rows = []

    def add_row(self, row):
        self.rows.append(row)
