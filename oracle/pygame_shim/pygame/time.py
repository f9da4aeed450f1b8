class Clock:
    def tick(self, fps=0):
        return 0
