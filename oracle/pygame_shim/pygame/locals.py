# pygame.locals: the reference does `from pygame.locals import *` and uses none of it
