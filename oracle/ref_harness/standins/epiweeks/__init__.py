class Week:  # stand-in: imported by models/covid/calibration/utils, not on the SEIRM path
    def __init__(self, year, week):
        self.year, self.week = year, week

    @classmethod
    def fromstring(cls, s):
        s = str(s)
        return cls(int(s[:4]), int(s[4:6]))

    def __add__(self, n):
        return Week(self.year, self.week + n)

    def __sub__(self, n):
        return Week(self.year, self.week - n)

    def cdcformat(self):
        return f"{self.year}{self.week:02d}"
